// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// Registration shim for the *unmodified* reference spconv 1.0 sources that live under
// /root/reference/modules/lib/spconv (they are compiled from where they lie by oracle/Makefile;
// nothing is copied into this repository).  The reference's own registration unit
// (src/spconv/all.cc:19-34) uses torch::jit::RegisterOperators(name, fn), an overload that no
// longer exists in torch 2.11, so the three operators on the 3D-WSIS hot path are registered
// here with TORCH_LIBRARY instead.  The operator names and the C++ entry points are the
// reference's: spconv_ops.h:27-33 (getIndicePair), :253-256 (indiceConv), :351-355
// (indiceConvBackward).
#include <spconv/spconv_ops.h>
#include <torch/library.h>

TORCH_LIBRARY(spconv, m) {
  m.def("get_indice_pairs_2d", &spconv::getIndicePair<2>);
  m.def("get_indice_pairs_3d", &spconv::getIndicePair<3>);
  m.def("indice_conv_fp32", &spconv::indiceConv<float>);
  m.def("indice_conv_backward_fp32", &spconv::indiceConvBackward<float>);
}
