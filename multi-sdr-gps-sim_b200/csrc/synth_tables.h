// Constant data of the GPS L1 C/A path, shared by host and device code.
#pragma once
#include <stdint.h>

namespace gpsb200 {

// First quadrant of the reference's 512-entry sine table (gps.c:145-178):
// floor(250*sin(2*pi*(k+0.5)/512)+0.5) for k = 0..127, except k = 35 where the
// reference holds 105 (formula: 105.50007). Remaining entries by symmetry:
// sin[255-k] = sin[k], sin[k+256] = -sin[k]; cos[k] = sin[(k+128)&511] (gps.c:180-213).
// tests/test_oracle.py / test_host_api.py check the expansion against the reference arrays.
#define GPSB200_QUARTER_SINE                                                              \
    2, 5, 8, 11, 14, 17, 20, 23, 26, 29, 32, 35, 38, 41, 44, 47,                          \
    50, 53, 56, 59, 62, 65, 68, 71, 74, 77, 80, 83, 86, 89, 91, 94,                       \
    97, 100, 103, 105, 108, 111, 114, 116, 119, 122, 125, 127, 130, 132, 135, 138,        \
    140, 143, 145, 148, 150, 153, 155, 157, 160, 162, 164, 167, 169, 171, 173, 176,       \
    178, 180, 182, 184, 186, 188, 190, 192, 194, 196, 198, 200, 202, 204, 205, 207,       \
    209, 210, 212, 214, 215, 217, 218, 220, 221, 223, 224, 225, 227, 228, 229, 230,       \
    232, 233, 234, 235, 236, 237, 238, 239, 240, 241, 241, 242, 243, 244, 244, 245,       \
    245, 246, 247, 247, 248, 248, 248, 249, 249, 249, 249, 250, 250, 250, 250, 250

// G2 code-phase delay in chips for PRN 1..32 (IS-GPS-200 Table 3-Ia; gps.c:273-278).
#define GPSB200_G2_DELAY                                                                  \
    5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258,                 \
    469, 470, 471, 472, 473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862

// C/A Gold code chips (0/1) of one PRN: G1 = x^10+x^3+1, G2 = x^10+x^9+x^8+x^6+x^3+x^2+1,
// both registers all ones, G2 delayed per PRN (gps.c:272-309).
inline int ca_code(int prn, uint8_t *ca /*[1023]*/) {
    static const uint16_t delay[32] = {GPSB200_G2_DELAY};
    if (prn < 1 || prn > 32) return -1;
    uint8_t g1[1023], g2[1023];
    unsigned r1 = 0x3FF, r2 = 0x3FF;
    for (int i = 0; i < 1023; i++) {
        g1[i] = (r1 >> 9) & 1;
        g2[i] = (r2 >> 9) & 1;
        unsigned f1 = ((r1 >> 2) ^ (r1 >> 9)) & 1;
        unsigned f2 = ((r2 >> 1) ^ (r2 >> 2) ^ (r2 >> 5) ^ (r2 >> 7) ^ (r2 >> 8) ^ (r2 >> 9)) & 1;
        r1 = ((r1 << 1) | f1) & 0x3FF;
        r2 = ((r2 << 1) | f2) & 0x3FF;
    }
    const int d = delay[prn - 1];
    for (int i = 0; i < 1023; i++) ca[i] = g1[i] ^ g2[(i + 1023 - d) % 1023];
    return 0;
}

}  // namespace gpsb200
