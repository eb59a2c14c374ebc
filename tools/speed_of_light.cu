// Speed-of-light probes for the per-sample synthesis formulation (measurement tool, not product).
//
// k_synth (csrc/synth_kernels.cu) is bound by instruction issue, not by HBM. These stripped kernels
// measure how far it is from what the SAME formulation could reach on this GPU if everything that
// is not strictly per channel-sample were free:
//
//   sol_quiet : exactly the "quiet path" of k_synth per 32-channel sample step (2 DADD, 2 DADD.RZ,
//               SHF, LOP3, IMAD, LDS, REDUX, UR->R move, one STS.128 per 4 samples) + the real
//               quantise/pack/store of every 64-sample chunk -- but no wrap prediction, no wrap
//               handling, no chip-window refill, no table build. = k_synth if no NCO ever wrapped.
//   sol_nostore : the same step stream without the quantise/pack/store of the chunks (one checksum word
//               per run is written instead): what the HBM stores themselves cost.
//
// Same launch shape as k_synth at 32 channels (21 warps per CTA, 2 CTAs per SM, 65.7 KB carrier table
// in shared memory), same bytes written per sample (2 B, int8 I/Q). Inputs are chosen so that no NCO
// leaves its range during a run (tiny carrier increments; the code phase is only used as a shift
// count), i.e. the instruction streams are the real ones but the results are not GPS signals.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -o gpsb200-sol tools/speed_of_light.cu
//   ./gpsb200-sol [blocks=2999] [reps=5]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {

constexpr int kBlockSamples = 300000;
constexpr int kRunSamples = 2400;
constexpr int kRuns = kBlockSamples / kRunSamples;   // 125
constexpr int kCtasPerBlock = 6;
constexpr int kWarps = 21;                           // runs per CTA
constexpr int kRows = 513;

struct Smem {
    int32_t atab[kRows][32];
    alignas(16) int32_t stage[24][64];
};

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));              \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

template <bool NOSTORE>
__global__ void __launch_bounds__(768, 2) k_sol(uint32_t *out, const int32_t *table, int nblk) {
    extern __shared__ __align__(16) unsigned char raw[];
    Smem &sm = *reinterpret_cast<Smem *>(raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kRows * 32; i += blockDim.x) (&sm.atab[0][0])[i] = table[i];
    __syncthreads();
    const int b = blockIdx.x / kCtasPerBlock, g = blockIdx.x - b * kCtasPerBlock;
    const int r = g * kWarps + warp;
    if (r >= kRuns) return;
    // per-lane NCO state: never wraps within 2400 steps
    double x = 0.001 * lane + 1e-4 * (r & 7), cc = 1.0e-6 * (lane + 1);
    double y = 3.0 + lane, dd = 0.341 + 1e-6 * lane;
    const double K43 = 8796093022208.0, K52 = 4503599627370496.0;
    const double KY = K52 - 3.0;
    const uint32_t w8 = 0xA5C3F096u ^ (lane * 0x9E3779B9u);
    const uint32_t abase = (uint32_t) __cvta_generic_to_shared(&sm.atab[0][lane]);
    int32_t *stage = &sm.stage[warp][0];
    const size_t samp0 = (size_t) b * kBlockSamples + (size_t) r * kRunSamples;
    int acc = 0;
    for (int s0 = 0; s0 < kRunSamples; s0 += 64) {
        const int len = kRunSamples - s0 >= 64 ? 64 : 32;
#pragma unroll 1
        for (int g8 = 0; g8 < len; g8 += 8) {
#pragma unroll
            for (int h = 0; h < 8; h += 4) {
                int sv[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int k = __double2loint(__dadd_rz(x, K43));
                    const int rel = __double2loint(__dadd_rz(y, KY));
                    const int kk = k ^ ((w8 >> (rel & 31)) & 0x100);
                    int e;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(e) : "r"(abase + (uint32_t) kk * 128u));
                    sv[i] = __reduce_add_sync(0xFFFFFFFFu, e);
                    x = __dadd_rn(x, cc);
                    y = __dadd_rn(y, dd);
                }
                *reinterpret_cast<int4 *>(&stage[g8 + h]) = make_int4(sv[0], sv[1], sv[2], sv[3]);
            }
        }
        __syncwarp();
        if (NOSTORE) {
            acc += stage[lane] + stage[lane + 32 < len ? lane + 32 : lane];
            __syncwarp();
            continue;
        }
        // quantise + pack + store, as k_synth's flush for int8
        if (len == 64) {
            uint32_t w = 0;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int p = stage[lane * 2 + t];
                const int iv = (int) (short) (p & 0xFFFF), qv = (p - iv) >> 16;
                const uint32_t two = (((uint32_t) (iv >> 4)) & 0xFFu) | ((((uint32_t) (qv >> 4)) & 0xFFu) << 8);
                w |= two << (16 * t);
            }
            out[(samp0 + s0) / 2 + lane] = w;
        } else {
            const int p = stage[lane];
            const int iv = (int) (short) (p & 0xFFFF), qv = (p - iv) >> 16;
            reinterpret_cast<uint16_t *>(out)[samp0 + s0 + lane] =
                (uint16_t) ((((uint32_t) (iv >> 4)) & 0xFFu) | ((((uint32_t) (qv >> 4)) & 0xFFu) << 8));
        }
        __syncwarp();
    }
    if (NOSTORE) out[(samp0 / 2) + lane] = (uint32_t) acc;
}


// ---------------------------------------------------------------------------------------------------
// sol_lanes: the instruction stream of a DIFFERENT formulation: LANE = SAMPLE, the
// channels are a loop. Both NCOs are evaluated as certified fixed-point LINEAR phases anchored at exact
// run-start states (the FP64 recurrence deviates from the exact linear phase by <= n * 2^-54 per step, so
// floor(512 * phase) taken from a 32-bit linear accumulator is the reference's index unless the fraction
// lies within a narrow band below an index boundary -- detected per sample, repaired on a slow path):
//   per channel and 96-sample window: one broadcast LDS.64 (window base phase, per-sample increment) and
//   one IMAD (this lane's phase); per channel-sample: SHF (index), IMAD + VIMNMX3/2 (band detection),
//   PRMT (sign mask of this channel from the lane's transposed sign word), LDS (table), IADD3/2 + IMAD
//   (sum and signed sum), IMAD (advance by 32 samples). Per window a lane=channel "prep" phase advances
//   the 64-bit phases, builds the three residue-class chip-sign words (sample 3q+r -> chip q + J (+1 after
//   one carry)), transposes them (5 shuffle stages each) and hands every lane its sample's sign word.
// Stand-in data; the prep arithmetic is representative in cost, not a GPS signal.
struct SmemLanes {
    int32_t tab[32][512];                 // [channel][k]: I + (Q << 16)
    uint2 base[16][32];                   // per warp: (window base phase, per-sample increment) per channel
    alignas(16) uint16_t stage[16][2][480];   // per warp, double buffered: 5 windows of 96 int8 I/Q pairs
};

__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
    // 32x32 bit-matrix transpose across the warp: lane i holds row i; afterwards lane j holds column j
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, x, s);
        x = (lane & s) ? ((x & ~m) | ((y >> s) & m)) : ((x & m) | ((y << s) & ~m));
    }
    return x;
}

template <bool DETECT, bool TMA>
__global__ void __launch_bounds__(512, 1) k_sol_lanes(uint32_t *out, const int32_t *table, int nblk, int nchan, unsigned *flags) {
    extern __shared__ __align__(16) unsigned char raw[];
    SmemLanes &sm = *reinterpret_cast<SmemLanes *>(raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 32 * 512; i += blockDim.x) (&sm.tab[0][0])[i] = table[i];
    __syncthreads();
    const uint32_t tbase = (uint32_t) __cvta_generic_to_shared(&sm.tab[0][0]);
    // persistent CTA: runs of 2400 samples, 125 per block, handed out round-robin to the warps of the grid
    const int nwarps = gridDim.x * 16;
    const long total_runs = (long) nblk * kRuns;
    for (long run = (long) blockIdx.x * 16 + warp; run < total_runs; run += nwarps) {
        // ---- lane = channel state ------------------------------------------------------------
        uint64_t P64 = 0x1234567890ABCDEFull * (lane + 1) + run * 0x9E3779B97F4A7C15ull;
        const uint64_t D64 = (uint64_t) ((0.0003 + 0.00005 * lane) * 18446744073709551616.0);   // cycles/sample
        uint64_t Y64 = ((uint64_t) (lane * 29 + 7) << 54) + (run & 1023) * 0x3FFFFFFFFFFull;   // chips << 54
        const uint64_t E64 = (uint64_t) (0.341 * 18014398509481984.0) + lane * 1000003ull;
        const double rinv = 1.0 / (double) (3 * E64 - (1ull << 54));
        uint32_t chips_lo = 0xA5C3F096u ^ (lane * 0x9E3779B9u), chips_hi = 0x3C96A5F0u + lane;
        const size_t samp0 = (size_t) run * kRunSamples;
        unsigned dmin_all = 0xFFFFFFFFu;
        int buf = 0;
        for (int w = 0; w < kRunSamples / 96; w++) {
            // ---- prep (lane = channel) ----------------------------------------------------
            sm.base[warp][lane] = make_uint2((uint32_t) (P64 >> 32) - 1u, (uint32_t) (D64 >> 32));
            uint32_t S[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const uint64_t phi = Y64 + (uint64_t) r * E64;
                const uint64_t F = phi & ((1ull << 54) - 1);
                const int J = (int) (phi >> 54) & 1;
                const double t = (double) ((1ull << 54) - F) * rinv;
                int q = __double2int_ru(t);
                q = q > 32 ? 32 : q;
                const uint32_t lowm = q >= 32 ? 0xFFFFFFFFu : ((1u << q) - 1u);
                const uint32_t a = __funnelshift_r(chips_lo, chips_hi, J), b2 = __funnelshift_r(chips_lo, chips_hi, J + 1);
                S[r] = (a & lowm) | (b2 & ~lowm);
                if (DETECT && (t - (double) (q - 1) < 1e-9)) dmin_all = 0;       // stand-in of the code band test
            }
            P64 += 96 * D64;
            Y64 += 96 * E64;
            chips_lo = __funnelshift_r(chips_lo, chips_hi, 9) * 0x9E3779B1u;     // stand-in of the window refill
            chips_hi ^= chips_lo >> 7;
            uint32_t Wt[3];
#pragma unroll
            for (int r = 0; r < 3; r++) Wt[r] = transpose32(S[r], lane);       // lane q: bit c = sign of channel c at sample 3q+r
            __syncwarp();
            // ---- main (lane = sample): three steps of 32 consecutive samples -------------------
            int accA[3] = {0, 0, 0}, accB[3] = {0, 0, 0};
            uint32_t Wsh[3][8];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int n = 32 * j + lane, q = n / 3, r = n - 3 * q;
                const uint32_t w0 = __shfl_sync(0xFFFFFFFFu, Wt[0], q), w1 = __shfl_sync(0xFFFFFFFFu, Wt[1], q),
                               w2 = __shfl_sync(0xFFFFFFFFu, Wt[2], q);
                const uint32_t ws = r == 0 ? w0 : (r == 1 ? w1 : w2);
#pragma unroll
                for (int i = 0; i < 8; i++) Wsh[j][i] = ws << (7 - i);        // channel c: byte c/8 of word c%8, sign bit on top
            }
            unsigned dmin = 0xFFFFFFFFu;
#pragma unroll 8
            for (int c = 0; c < 32; c += 2) {
                if (c >= nchan) break;
                const uint2 b0 = sm.base[warp][c], b1 = sm.base[warp][c + 1];
                uint32_t P0 = b0.x + (uint32_t) lane * b0.y, P1 = b1.x + (uint32_t) lane * b1.y;
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const uint32_t k0 = P0 >> 23, k1 = P1 >> 23;
                    if (DETECT) dmin = __vimin3_u32(dmin, P0 * 0xFFFFFE00u - 512u, P1 * 0xFFFFFE00u - 512u);
                    const int M0 = (int) __byte_perm(Wsh[j][c & 7], 0, 0x8888 | ((c >> 3) * 0x1111));
                    const int M1 = (int) __byte_perm(Wsh[j][(c + 1) & 7], 0, 0x8888 | (((c + 1) >> 3) * 0x1111));
                    int e0, e1;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(e0) : "r"(tbase + c * 2048 + k0 * 4));
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(e1) : "r"(tbase + (c + 1) * 2048 + k1 * 4));
                    accA[j] += e0 + e1;
                    accB[j] += e0 * M0;
                    accB[j] += e1 * M1;
                    P0 += 32u * b0.y;
                    P1 += 32u * b1.y;
                }
            }
            dmin_all = min(dmin_all, dmin);
            // ---- quantise + pack + store ------------------------------------------------------------
            const int slot = w % 5;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int p = accA[j] + 2 * accB[j];
                const int iv = (int) (short) (p & 0xFFFF), qv = (p - iv) >> 16;
                const uint16_t two = (uint16_t) ((((uint32_t) (iv >> 4)) & 0xFFu) | ((((uint32_t) (qv >> 4)) & 0xFFu) << 8));
                if (TMA) sm.stage[warp][buf][slot * 96 + 32 * j + lane] = two;
                else reinterpret_cast<uint16_t *>(out)[samp0 + (size_t) w * 96 + 32 * j + lane] = two;
            }
            if (TMA && slot == 4) {
                // 480 samples (960 bytes) of this warp go out as ONE bulk copy shared -> global
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    const uint32_t src = (uint32_t) __cvta_generic_to_shared(&sm.stage[warp][buf][0]);
                    char *dst = reinterpret_cast<char *>(out) + (samp0 + (size_t) (w - 4) * 96) * 2;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(960) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");    // the OTHER buffer is free again
                }
                __syncwarp();
                buf ^= 1;
            }
        }
        if (TMA && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (DETECT && dmin_all < 65536u) atomicAdd(flags, 1u);
    }
}

template <bool DETECT, bool TMA>
float run_lanes(uint32_t *out, const int32_t *tab, int nblk, int nchan, int reps, unsigned *flags) {
    CK(cudaFuncSetAttribute(k_sol_lanes<DETECT, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(SmemLanes)));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; i++) k_sol_lanes<DETECT, TMA><<<sms, 512, sizeof(SmemLanes)>>>(out, tab, nblk, nchan, flags);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) k_sol_lanes<DETECT, TMA><<<sms, 512, sizeof(SmemLanes)>>>(out, tab, nblk, nchan, flags);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

// ---------------------------------------------------------------------------------------------------
// sol_half_table: k_synth's quiet path with the HALF carrier table (rows k and k + 256 are negatives, so
// 256 rows = 32.8 KB suffice) -- the sign (chip x data bit x table half) is applied to the looked-up value
// instead of being folded into the row index. Buys a third resident CTA per SM (63 instead of 42 warps),
// costs instructions per channel-sample: the question is which effect wins on an issue-bound kernel.
struct SmemHalf {
    int32_t atab[256][32];
    alignas(16) int32_t stage[24][64];
};

__global__ void __launch_bounds__(672, 3) k_sol_half(uint32_t *out, const int32_t *table, int nblk) {
    extern __shared__ __align__(16) unsigned char raw[];
    SmemHalf &sm = *reinterpret_cast<SmemHalf *>(raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256 * 32; i += blockDim.x) (&sm.atab[0][0])[i] = table[i];
    __syncthreads();
    const int b = blockIdx.x / kCtasPerBlock, g = blockIdx.x - b * kCtasPerBlock;
    const int r = g * kWarps + warp;
    if (r >= kRuns) return;
    double x = 0.001 * lane + 1e-4 * (r & 7), cc = 1.0e-6 * (lane + 1);
    double y = 3.0 + lane, dd = 0.341 + 1e-6 * lane;
    const double K43 = 8796093022208.0, K52 = 4503599627370496.0;
    const double KY = K52 - 3.0;
    const uint32_t w8 = 0xA5C3F096u ^ (lane * 0x9E3779B9u);
    const uint32_t abase = (uint32_t) __cvta_generic_to_shared(&sm.atab[0][lane]);
    int32_t *stage = &sm.stage[warp][0];
    const size_t samp0 = (size_t) b * kBlockSamples + (size_t) r * kRunSamples;
    for (int s0 = 0; s0 < kRunSamples; s0 += 64) {
        const int len = kRunSamples - s0 >= 64 ? 64 : 32;
#pragma unroll 1
        for (int g8 = 0; g8 < len; g8 += 8) {
#pragma unroll
            for (int h = 0; h < 8; h += 4) {
                int sv[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int k = __double2loint(__dadd_rz(x, K43));
                    const int rel = __double2loint(__dadd_rz(y, KY));
                    const int kk = k ^ ((w8 >> (rel & 31)) & 0x100);
                    int e;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(e) : "r"(abase + (uint32_t) (kk & 255) * 128u));
                    const int m = -((kk >> 8) & 1);                // 0 or -1: negate the packed I + (Q << 16)
                    e = (e ^ m) - m;
                    sv[i] = __reduce_add_sync(0xFFFFFFFFu, e);
                    x = __dadd_rn(x, cc);
                    y = __dadd_rn(y, dd);
                }
                *reinterpret_cast<int4 *>(&stage[g8 + h]) = make_int4(sv[0], sv[1], sv[2], sv[3]);
            }
        }
        __syncwarp();
        if (len == 64) {
            uint32_t w = 0;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int p = stage[lane * 2 + t];
                const int iv = (int) (short) (p & 0xFFFF), qv = (p - iv) >> 16;
                const uint32_t two = (((uint32_t) (iv >> 4)) & 0xFFu) | ((((uint32_t) (qv >> 4)) & 0xFFu) << 8);
                w |= two << (16 * t);
            }
            out[(samp0 + s0) / 2 + lane] = w;
        } else {
            const int p = stage[lane];
            const int iv = (int) (short) (p & 0xFFFF), qv = (p - iv) >> 16;
            reinterpret_cast<uint16_t *>(out)[samp0 + s0 + lane] =
                (uint16_t) ((((uint32_t) (iv >> 4)) & 0xFFu) | ((((uint32_t) (qv >> 4)) & 0xFFu) << 8));
        }
        __syncwarp();
    }
}

float run_half(uint32_t *out, const int32_t *tab, int nblk, int reps) {
    CK(cudaFuncSetAttribute(k_sol_half, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(SmemHalf)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) k_sol_half<<<nblk * kCtasPerBlock, kWarps * 32, sizeof(SmemHalf)>>>(out, tab, nblk);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) k_sol_half<<<nblk * kCtasPerBlock, kWarps * 32, sizeof(SmemHalf)>>>(out, tab, nblk);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sol_half, kWarps * 32, sizeof(SmemHalf));
    printf("{\"formulation\": \"k_synth quiet path, 256-row table, sign applied to the value\", \"blocks\": %d, \"half_table_ms\": %.3f, "
           "\"ctas_per_sm\": %d}\n", nblk, ms / reps, occ);
    return ms / reps;
}

template <bool F>
float run(uint32_t *out, const int32_t *tab, int nblk, int reps) {
    CK(cudaFuncSetAttribute(k_sol<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(Smem)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) k_sol<F><<<nblk * kCtasPerBlock, kWarps * 32, sizeof(Smem)>>>(out, tab, nblk);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) k_sol<F><<<nblk * kCtasPerBlock, kWarps * 32, sizeof(Smem)>>>(out, tab, nblk);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

}  // namespace

int main(int argc, char **argv) {
    const int nblk = argc > 1 ? atoi(argv[1]) : 2999;
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    const double peak = argc > 3 ? atof(argv[3]) : 6586.1;     // GB/s, MEASURED_PEAKS.json of this pool
    std::vector<int32_t> tab((size_t) kRows * 32);
    for (size_t i = 0; i < tab.size(); i++) tab[i] = (int) ((i * 2654435761u) % 401) - 200 + (((int) ((i * 40503u) % 401) - 200) << 16);
    int32_t *d_tab;
    uint32_t *d_out;
    CK(cudaMalloc(&d_tab, tab.size() * 4));
    CK(cudaMemcpy(d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, (size_t) nblk * kBlockSamples * 2));
    const double samples = (double) nblk * kBlockSamples;
    run_half(d_out, d_tab, nblk, reps);
    const float q = run<false>(d_out, d_tab, nblk, reps), f = run<true>(d_out, d_tab, nblk, reps);
    {
        std::vector<int32_t> t2((size_t) 32 * 512);
        for (size_t i = 0; i < t2.size(); i++) t2[i] = (int) ((i * 2654435761u) % 401) - 200 + (((int) ((i * 40503u) % 401) - 200) << 16);
        int32_t *d_t2;
        unsigned *d_flags;
        CK(cudaMalloc(&d_t2, t2.size() * 4));
        CK(cudaMemcpy(d_t2, t2.data(), t2.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_flags, 4));
        CK(cudaMemset(d_flags, 0, 4));
        const float a = run_lanes<true, false>(d_out, d_t2, nblk, 32, reps, d_flags);
        const float b = run_lanes<false, false>(d_out, d_t2, nblk, 32, reps, d_flags);
        const float c = run_lanes<true, true>(d_out, d_t2, nblk, 32, reps, d_flags);
        const float d12 = run_lanes<true, true>(d_out, d_t2, nblk, 12, reps, d_flags);
        unsigned fl = 0;
        CK(cudaMemcpy(&fl, d_flags, 4, cudaMemcpyDeviceToHost));
        printf("{\"formulation\": \"lanes=samples, certified fixed-point linear phases\", \"blocks\": %d, "
               "\"lanes_detect_stg_ms\": %.3f, \"lanes_nodetect_stg_ms\": %.3f, \"lanes_detect_tma_store_ms\": %.3f, "
               "\"lanes_detect_tma_store_12ch_ms\": %.3f, \"lanes_gsps\": %.2f, \"lanes_hbm_frac\": %.4f, \"flagged_runs\": %u}\n",
               nblk, a, b, c, d12, samples / c / 1e6, samples * 2 / (c * 1e-3) / 1e9 / peak, fl);
    }
    printf("{\"blocks\": %d, \"channels\": 32, \"sol_quiet_ms\": %.3f, \"sol_quiet_gsps\": %.2f, \"sol_quiet_hbm_frac\": %.4f, "
           "\"sol_nostore_ms\": %.3f, \"sol_nostore_gsps\": %.2f, \"sol_nostore_hbm_frac\": %.4f, \"hbm_peak_gbs\": %.1f}\n",
           nblk, q, samples / q / 1e6, samples * 2 / (q * 1e-3) / 1e9 / peak, f, samples / f / 1e6,
           samples * 2 / (f * 1e-3) / 1e9 / peak, peak);
    return 0;
}
