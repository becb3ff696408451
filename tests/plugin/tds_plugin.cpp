// Architecture plugin in the style of recipes/slimIPL/100h_supervised.cpp: a Module subclass plus
//   extern "C" fl::Module* createModule(int64_t nFeature, int64_t nLabel)
// built against include/fl_compat/fl_compat.h and linked to libw2l_b200.so.  Used by tests/test_plugin.py.
#include "fl_compat/fl_compat.h"

namespace {
const char* kArch =
    "V -1 NFEAT 1 0\n"
    "C2 1 4 5 1 2 1 -1 -1\n"
    "R\n"
    "LN 3\n"
    "TDS 4 5 80 0.0\n"
    "V 0 320 1 0\n"
    "RO 1 0 3 2\n"
    "L 320 NLABEL\n";

class SmallTds : public fl::Module {
 public:
  SmallTds(int64_t nFeature, int64_t nLabel) : body_(fl::pkg::runtime::buildSequentialModule(kArch, nFeature, nLabel)) {}
  std::vector<fl::Variable> forward(const std::vector<fl::Variable>& input) override {
    return body_->forward(std::vector<fl::Variable>{input[0]});  // input[1] (sizes) is unused by this arch
  }
  std::vector<fl::Variable> params() const override { return body_->params(); }
  void setParams(const fl::Variable& v, int i) override { body_->setParams(v, i); }
  void train() override {
    train_ = true;
    body_->train();
  }
  void eval() override {
    train_ = false;
    body_->eval();
  }
  std::string prettyString() const override { return "SmallTds plugin\n" + body_->prettyString(); }

 private:
  std::shared_ptr<fl::Sequential> body_;
};
}  // namespace

extern "C" __attribute__((visibility("default"))) fl::Module* createModule(int64_t nFeature, int64_t nLabel) {
  return new SmallTds(nFeature, nLabel);
}
