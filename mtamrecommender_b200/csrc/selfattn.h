// Self-attention model family (PISTRec / SASRec / TA-SASRec / TiSASRec): layout + workspace hooks.
#pragma once
#include <stddef.h>
#include <string>
#include <vector>

#include "../../include/mtam.h"

namespace mtam {
struct ParamDesc;
int sa_build_layout(const mtam_config& c, size_t& offset, std::vector<ParamDesc>& params, size_t& lnfb, size_t& lnfg);
size_t sa_workspace_bytes(const mtam_config& c);
}  // namespace mtam
