// Exact fast-forward of the reference's double-precision NCO recurrences.
//
// The reference advances two phases once per sample, in IEEE binary64 with a
// separately rounded add (no FMA; -std=c11 => -ffp-contract=off):
//   code   : x += c; if (x >= 1023.0) { x -= 1023.0; ++periods; }      (gps.c:2789-2793)
//   carrier: x += c; if (x >= 1.0) x -= 1.0; else if (x < 0.0) x += 1.0; (gps.c:2821-2826)
// with c = fl(f * delt) constant inside a 0.1 s block. Bit-exact output needs the
// exact x after n steps, not x0 + n*c (SURVEY.md hard part 1).
//
// Observation used here: while x stays inside one binade [2^e, 2^(e+1)) and does not
// wrap, x = X*u (u = 2^(e-52), X a 53-bit integer) and fl(x + c) = (X + R)*u with a
// CONSTANT integer R = round-to-nearest(c/u) (on an exact tie: the even one of the two
// neighbours, once X is even). So a run of k in-binade steps is X += k*R, and only the
// steps that cross a binade boundary or wrap are executed as real additions.
// nco_advance() therefore costs O(#binade crossings) instead of O(n).
//
// Shared by host code (carrier-phase chain across blocks) and device code (run
// checkpoints), hence the GPSB_HD qualifier and the absence of libm calls.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GPSB_HD __host__ __device__ __forceinline__
#else
#define GPSB_HD inline
#endif

namespace gpsb200 {

enum NcoKind { NCO_CODE = 0, NCO_CARRIER = 1 };

// Largest double below 1.0. The reference's negative-Doppler wrap `carr_phase += 1.0`
// (gps.c:2825-2826) can round a tiny negative phase to exactly 1.0, after which its table
// index (int) floor(1.0 * 512) = 512 reads past cosTable512/sinTable512 (gps.c:2775-2782):
// undefined behaviour, probability ~2^-54/|c| per wrap. This implementation (and the test
// oracle) clamp that one value to 1 - 2^-53 instead, which keeps the index at 511.
constexpr double kBelowOne = 0.99999999999999988897769753748434595763683319091796875;

GPSB_HD uint64_t f64_bits(double v) {
#if defined(__CUDA_ARCH__)
    return (uint64_t) __double_as_longlong(v);
#else
    uint64_t b;
    memcpy(&b, &v, 8);
    return b;
#endif
}

GPSB_HD double bits_f64(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long) b);
#else
    double v;
    memcpy(&v, &b, 8);
    return v;
#endif
}

// One step exactly as the reference performs it. `periods` counts code wraps.
template <int KIND>
GPSB_HD void nco_step(double &x, double c, int64_t &periods) {
#if defined(__CUDA_ARCH__)
    x = __dadd_rn(x, c);
#else
    x = x + c;
#endif
    if (KIND == NCO_CODE) {
        if (x >= 1023.0) {
            x -= 1023.0;
            ++periods;
        }
    } else {
        if (x >= 1.0) x -= 1.0;
        else if (x < 0.0) {
            x += 1.0;
            if (x >= 1.0) x = kBelowOne;
        }
    }
}

// Advance x by exactly n reference steps of increment c.
// Returns the number of loop iterations spent (diagnostics only).
template <int KIND>
GPSB_HD int nco_advance(double &x, double c, int64_t n, int64_t &periods) {
    const uint64_t MANT = (1ull << 52) - 1;
    int iters = 0;
    if (c == 0.0 || n <= 0) return 0;
    const bool neg = c < 0.0;
    const uint64_t cb = f64_bits(c) & 0x7FFFFFFFFFFFFFFFull;   // |c|
    const int ec = (int) (cb >> 52);
    const uint64_t cm = (cb & MANT) | (1ull << 52);
    const double ac = bits_f64(cb);
    // Mantissa-domain upper bound of the wrap threshold in its own binade:
    // 1023.0 = 0x3FF<<... lives in binade [512,1024): 1023 * 2^43; 1.0 is a binade edge.
    const int e_wrap = (KIND == NCO_CODE) ? (1023 + 9) : -1;
    const uint64_t x_wrap = 1023ull << 43;

    while (n > 0) {
        ++iters;
        const uint64_t xb = f64_bits(x);
        const int ex = (int) ((xb >> 52) & 0x7FF);
        // Below |c| (incl. zero, negatives, subnormals), or c subnormal: plain step.
        if ((xb >> 63) || !(x >= ac) || ec == 0) {
            nco_step<KIND>(x, c, periods);
            --n;
            continue;
        }
        const int sh = ex - ec;                  // >= 0 because x >= |c|
        if (sh >= 54) return iters;              // |c| < ulp(x)/2: x + c == x for ever
        uint64_t X = (xb & MANT) | (1ull << 52);
        uint64_t Cq = cm >> sh;
        uint64_t R = Cq;
        if (sh > 0) {
            const uint64_t rem = cm & ((1ull << sh) - 1);
            const uint64_t half = 1ull << (sh - 1);
            if (rem > half) R = Cq + 1;
            else if (rem == half) {              // exact tie: result mantissa is even
                if (X & 1) {                     // one real step makes X even
                    nco_step<KIND>(x, c, periods);
                    --n;
                    continue;
                }
                R = (Cq & 1) ? Cq + 1 : Cq;
            }
        }
        if (R == 0) return iters;                // increment rounds away: fixed point
        // Steps that provably stay strictly inside the binade (and below the wrap
        // threshold); Cq+1 >= R keeps every intermediate REAL sum inside as well.
        uint64_t room;
        if (!neg) {
            uint64_t hi = (KIND == NCO_CODE && ex == e_wrap) ? x_wrap : (1ull << 53);
            room = hi - 1 - X;
        } else {
            room = X - (1ull << 52);
        }
        uint64_t k = room / (Cq + 1);
        if (k > (uint64_t) n) k = (uint64_t) n;
        if (k > 0) {
            X = neg ? X - k * R : X + k * R;
            x = bits_f64(((uint64_t) ex << 52) | (X & MANT));
            n -= (int64_t) k;
        }
        if (n > 0) {                             // boundary / wrap (or filler) step, done for real
            nco_step<KIND>(x, c, periods);
            --n;
        }
    }
    return iters;
}

// NAV-message position after `periods` more code periods (gps.c:2793-2812):
// 20 code periods per data bit, 30 bits per word.
GPSB_HD void nav_advance(int &iword, int &ibit, int &icode, int64_t periods) {
    int64_t t = (int64_t) icode + periods;
    icode = (int) (t % 20);
    t = (int64_t) ibit + t / 20;
    ibit = (int) (t % 30);
    iword += (int) (t / 30);
}

}  // namespace gpsb200
