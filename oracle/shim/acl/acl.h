// TEST INFRASTRUCTURE ONLY -- stand-in for CANN's `acl/acl.h`; the reference's src/data_utils.h:16
// includes it only for aclError / ACL_ERROR_NONE (CHECK_ACL, :41-47) and aclFloat16 (:136-140).
#pragma once
#include <cstdint>
#include <cstring>
typedef int aclError;
#define ACL_ERROR_NONE 0
typedef uint16_t aclFloat16;
static inline float aclFloat16ToFloat(aclFloat16 h) {
    uint32_t sign = (uint32_t)(h >> 15) << 31, exp = (h >> 10) & 0x1f, man = h & 0x3ff, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {
            int e = -1;
            do {
                man <<= 1;
                e++;
            } while (!(man & 0x400));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112) << 23) | (man << 13);
    }
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}
