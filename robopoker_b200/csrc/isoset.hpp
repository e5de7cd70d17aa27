// isoset.hpp — the device-resident isomorphism list behind `rbp_isoset_t` (iso.cu builds it; nlhe.cu turns it into the
// blueprint solver's abstraction lookup).
#pragma once
#include <cstdint>

struct rbp_isoset {
    int street = 0, device = 0;
    int64_t n = 0;
    uint64_t* pocket = nullptr;  // canonical hole masks, enumeration order (sorted by (pocket, public))
    uint64_t* pub = nullptr;     // canonical board masks
    uint8_t* abs = nullptr;      // optional abstraction column (lookup table iso → bucket)
    bool have_abs = false;
};
