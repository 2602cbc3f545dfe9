// Self-attention model family (Tq = L blocks).  Not built yet in this revision.
#include "common.cuh"
#include "selfattn.h"

namespace mtam {
int sa_build_layout(const mtam_config& c, size_t& offset, std::vector<ParamDesc>& params, size_t& lnfb, size_t& lnfg) {
  (void)offset; (void)params; (void)lnfb; (void)lnfg;
  return set_error(MTAM_ERR_UNSUPPORTED, "model kind %d (self-attention family) is not built yet", c.kind);
}
size_t sa_workspace_bytes(const mtam_config&) { return 0; }
}  // namespace mtam
