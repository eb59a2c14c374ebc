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
    const float q = run<false>(d_out, d_tab, nblk, reps), f = run<true>(d_out, d_tab, nblk, reps);
    printf("{\"blocks\": %d, \"channels\": 32, \"sol_quiet_ms\": %.3f, \"sol_quiet_gsps\": %.2f, \"sol_quiet_hbm_frac\": %.4f, "
           "\"sol_nostore_ms\": %.3f, \"sol_nostore_gsps\": %.2f, \"sol_nostore_hbm_frac\": %.4f, \"hbm_peak_gbs\": %.1f}\n",
           nblk, q, samples / q / 1e6, samples * 2 / (q * 1e-3) / 1e9 / peak, f, samples / f / 1e6,
           samples * 2 / (f * 1e-3) / 1e9 / peak, peak);
    return 0;
}
