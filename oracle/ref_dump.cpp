// ref_dump.cpp - TEST INFRASTRUCTURE (parity oracle), never linked into the product.
//
// A small pybind11 module that instantiates the REFERENCE's own Metadata<3>
// (/root/reference/SparseConvNet/sparseconvnet/SCN/Metadata/Metadata.cpp, included where it
// lies - no reference source is copied into this repo) and exposes its integer state as
// tensors: input-layer rulebook, submanifold / strided / SparseToDense rulebooks, nActive and
// spatial locations.  The reference's pybind surface (pybind.cpp) does not export rulebooks;
// this is how the golden fixtures under tests/golden/ and SURVEY.md Appendix C were produced.
// Built by oracle/build_ref.py into oracle/_ref/SCN_refdump.so.
#define ENABLE_OPENMP YES
#include <omp.h>
#include <torch/extension.h>

#include "Metadata/Metadata.cpp"
template class Metadata<3>;

namespace {

at::Tensor to_tensor(const std::vector<Int> &v, long cols) {
  at::Tensor t = torch::zeros({(long)v.size() / cols, cols}, torch::kInt32);
  if (!v.empty()) std::memcpy(t.data_ptr<int32_t>(), v.data(), v.size() * sizeof(Int));
  return t;
}

std::vector<at::Tensor> rulebook_tensors(const RuleBook &rb) {
  std::vector<at::Tensor> out;
  for (auto const &r : rb) out.push_back(to_tensor(r, 2));
  return out;
}

struct RefMetadata3 {
  Metadata<3> m;

  void inputLayer(at::Tensor spatialSize, at::Tensor coords, long batchSize, long mode) {
    m.inputLayer(spatialSize, coords, (Int)batchSize, (Int)mode);
  }
  // [header(4 ints)], table nOut x (1+maxActive)   (IOLayersRules.h:10-15,112-124)
  std::vector<at::Tensor> inputLayerRuleBook() {
    std::vector<at::Tensor> out;
    auto &rb = m.inputLayerRuleBook;
    out.push_back(to_tensor(rb[0], (long)rb[0].size()));
    if (rb.size() > 1) {
      const long w = 1 + rb[0][1];
      out.push_back(to_tensor(rb[1], w));
    }
    return out;
  }
  std::vector<at::Tensor> getSubmanifoldRuleBook(at::Tensor spatialSize, at::Tensor size) {
    return rulebook_tensors(m.getSubmanifoldRuleBook(spatialSize, size, true));
  }
  std::vector<at::Tensor> getRuleBook(at::Tensor inSize, at::Tensor outSize, at::Tensor size,
                                      at::Tensor stride) {
    return rulebook_tensors(m.getRuleBook(inSize, outSize, size, stride, true));
  }
  std::vector<at::Tensor> getSparseToDenseRuleBook(at::Tensor spatialSize) {
    return rulebook_tensors(m.getSparseToDenseRuleBook(spatialSize, true));
  }
  long getNActive(at::Tensor spatialSize) { return m.getNActive(spatialSize); }
  at::Tensor getSpatialLocations(at::Tensor spatialSize) {
    return m.getSpatialLocations(spatialSize);
  }
};

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, mod) {
  pybind11::class_<RefMetadata3>(mod, "RefMetadata3")
      .def(pybind11::init<>())
      .def("inputLayer", &RefMetadata3::inputLayer)
      .def("inputLayerRuleBook", &RefMetadata3::inputLayerRuleBook)
      .def("getSubmanifoldRuleBook", &RefMetadata3::getSubmanifoldRuleBook)
      .def("getRuleBook", &RefMetadata3::getRuleBook)
      .def("getSparseToDenseRuleBook", &RefMetadata3::getSparseToDenseRuleBook)
      .def("getNActive", &RefMetadata3::getNActive)
      .def("getSpatialLocations", &RefMetadata3::getSpatialLocations);
}
