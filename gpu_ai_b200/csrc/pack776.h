// gpu_ai_b200/csrc/pack776.h -- reference `State` (776 B) -> b2p_state16 converter (pack776.cpp)
#pragma once

#include <stddef.h>

#include "../../include/b2p.h"

namespace b2p {

// n consecutive reference States -> n packed states (single thread; callers split the range over their workers)
void pack776_range(const unsigned char *states, size_t n, b2p_state16 *out);
void pack776_range_scalar(const unsigned char *states, size_t n, b2p_state16 *out);  // the portable path, for tests
const char *pack776_impl();  // "avx512bw+bmi2" | "scalar"

}  // namespace b2p
