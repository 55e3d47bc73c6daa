// shared between capi_ministep.cc (pure host) and capi_host.cc (LaplaceProblem hooks)
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "ministep.h"

namespace ministep_capi {
extern thread_local std::string g_err;
// dtype: 0 = int32, 1 = int64, 2 = float64, 3 = uint8, 4 = float32
int get_array(const ministep::Forest *fp, const ministep::DoFs *dp, const ministep::Csr *system,
              const ministep::LevelOperators *ops, const std::vector<std::vector<float>> *eta,
              const std::vector<std::vector<char>> *flags, const char *name, int l, const void **ptr, int64_t *count,
              int *dtype);
}  // namespace ministep_capi
