// Gather roofline micro-benchmark: how fast can a B200 gather whole fp32
// feature rows when there is NO sparse-matrix bookkeeping at all?
//
// The kernel mirrors spmm_rows_kernel's access pattern (one warp per output row, `deg` gathered
// rows of `F` floats each, 8 gathers in flight per warp, float4 lanes, accumulate, one streamed
// output row) but takes its column ids from a hash instead of a CSR, so what is measured is the
// memory system: L1 / L2->SM fabric / DRAM under row gathers of the products shape.
//
// Column distribution (per entry):  with probability p_uni the column is uniform over all rows
// (always a DRAM miss once H >> L2); otherwise |col - row| is log-uniform in [1, dmax] — the
// distance law of the synthetic products graph (gnn-tf_b200/synthetic.py:powerlaw_edges), so
// dmax = N/2 reproduces its locality, a small dmax gives an all-L1/L2-hit gather and p_uni = 1 an
// all-miss gather.
//
// Output: one JSON object per configuration on stdout.  Build: make -C scripts/microbench.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                            \
    do {                                                                                 \
        cudaError_t e_ = (x);                                                            \
        if (e_ != cudaSuccess) {                                                         \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                     \
        }                                                                                \
    } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// column of entry e of row `row`
__device__ __forceinline__ int pick_col(int row, int e, int n, float log2_dmax, uint32_t uni_thresh, uint32_t seed) {
    const uint32_t h = mix((uint32_t)row * 0x9e3779b9u + (uint32_t)e * 0x85ebca6bu + seed);
    const uint32_t h2 = mix(h + 0x68bc21ebu);
    if (h2 < uni_thresh) return (int)(((uint64_t)h * (uint64_t)n) >> 32);
    const float u = (float)(h >> 8) * (1.0f / 16777216.0f);
    int d = (int)exp2f(u * log2_dmax);
    d = max(1, d);
    int c = row + ((h2 & 1u) ? d : -d);
    if (c < 0) c += n;
    if (c >= n) c -= n;
    return c;
}

// LANES active lanes per warp own float4 slot `lane` of every gathered row (F = 4*LANES).
template <int UNROLL>
__global__ void __launch_bounds__(256, 5)
gather_ldg_kernel(const float* __restrict__ H, float* __restrict__ out, int n, int ld, int lanes, int deg,
                  float log2_dmax, uint32_t uni_thresh, uint32_t seed, int rows_per_cta, int write_out) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool on = lane < lanes;
    const float* Hl = H + (on ? lane * 4 : 0);
    const uint32_t pitch = (uint32_t)ld * 4u;
    for (int it = 0; it < rows_per_cta; it += 8) {
        const int row = blockIdx.x * rows_per_cta + it + warp;
        if (row >= n) break;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int base = 0; base < deg; base += 32) {
            const int mycol = pick_col(row, base + lane, n, log2_dmax, uni_thresh, seed);
            const int cnt = min(32, deg - base);
            for (int j = 0; j < cnt; j += UNROLL) {
                float4 x[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const int c = __shfl_sync(0xffffffffu, mycol, (j + u) & 31);
                    const float* p;
                    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(c), "r"(pitch), "l"(Hl));
                    if (on) x[u] = __ldg(reinterpret_cast<const float4*>(p));
                    else x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const float w = (j + u < cnt) ? 0.02f : 0.0f;
                    acc.x = fmaf(w, x[u].x, acc.x); acc.y = fmaf(w, x[u].y, acc.y);
                    acc.z = fmaf(w, x[u].z, acc.z); acc.w = fmaf(w, x[u].w, acc.w);
                }
            }
        }
        if (on) {
            if (write_out) {
                float* o = out + (size_t)row * ld + lane * 4;
                asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(acc.x), "f"(acc.y),
                             "f"(acc.z), "f"(acc.w)
                             : "memory");
            } else if (acc.x == 123.456f) {
                out[0] = acc.y;
            }
        }
    }
}

struct Cfg {
    const char* name;
    int n, F, ld, deg;
    double dmax;   // <= 0: N/2
    double p_uni;
    int write_out;
    int rows_per_cta;
};

int main(int argc, char** argv) {
    int reps = 5;
    const char* filter = nullptr;  // run only configurations whose name contains this substring
    if (argc > 1) reps = atoi(argv[1]);
    if (argc > 2) filter = argv[2];
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    fprintf(stderr, "device %s, %d SMs, L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);

    const int NP = 2449029;  // products rows
    std::vector<Cfg> cfgs = {
        // ---- ceilings
        {"l2_hit_small_H_F100", 80000, 100, 100, 50, 0, 1.0, 1, 64},        // H = 32 MB: every gather an L2 hit
        {"l2_hit_small_H_F128", 80000, 128, 128, 50, 0, 1.0, 1, 64},
        {"l1_hit_band16_F100", NP, 100, 100, 50, 16, 0.0, 1, 64},            // +-16 rows: L1-resident band
        {"l1l2_band256_F100", NP, 100, 100, 50, 256, 0.0, 1, 64},
        {"l2_band4k_F100", NP, 100, 100, 50, 4096, 0.0, 1, 64},
        {"l2_band32k_F100", NP, 100, 100, 50, 32768, 0.0, 1, 64},
        {"dram_miss_uniform_F100", NP, 100, 100, 50, 0, 1.0, 1, 64},         // H = 980 MB uniform: ~all DRAM misses
        {"dram_miss_uniform_F100_ld128", NP, 100, 128, 50, 0, 1.0, 1, 64},   // aligned 512-byte pitch
        {"dram_miss_uniform_F128", NP, 128, 128, 50, 0, 1.0, 1, 64},
        {"dram_miss_uniform_F100_nowrite", NP, 100, 100, 50, 0, 1.0, 0, 64},
        // ---- the products distance law (log-uniform up to N/2) and mixtures
        {"products_law_F100", NP, 100, 100, 50, -1, 0.0, 1, 64},
        {"products_law_F100_rpc8", NP, 100, 100, 50, -1, 0.0, 1, 8},
        {"products_law_F100_rpc512", NP, 100, 100, 50, -1, 0.0, 1, 512},
        {"products_law_F128", NP, 128, 128, 50, -1, 0.0, 1, 64},
        {"band32k_plus_10pct_uniform", NP, 100, 100, 50, 32768, 0.10, 1, 64},
        {"band32k_plus_20pct_uniform", NP, 100, 100, 50, 32768, 0.20, 1, 64},
        {"band32k_plus_35pct_uniform", NP, 100, 100, 50, 32768, 0.35, 1, 64},
        // ---- the arxiv shape (everything L2-resident)
        {"arxiv_law_F128", 169343, 128, 128, 14, -1, 0.0, 1, 8},
        {"arxiv_law_F40", 169343, 40, 40, 14, -1, 0.0, 1, 8},
    };

    size_t max_elems = 0;
    for (auto& c : cfgs) max_elems = std::max(max_elems, (size_t)c.n * c.ld);
    float *H, *out;
    CK(cudaMalloc(&H, max_elems * sizeof(float)));
    CK(cudaMalloc(&out, max_elems * sizeof(float)));
    CK(cudaMemset(H, 0, max_elems * sizeof(float)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    for (auto& c : cfgs) {
        if (filter && !strstr(c.name, filter)) continue;
        const double dmax = c.dmax > 0 ? c.dmax : c.n / 2.0;
        const float log2_dmax = (float)std::log2(dmax);
        const uint32_t thr = c.p_uni >= 1.0 ? 0xffffffffu : (uint32_t)(c.p_uni * 4294967296.0);
        const int lanes = c.F / 4;
        const int grid = (c.n + c.rows_per_cta - 1) / c.rows_per_cta;
        std::vector<float> ms;
        for (int r = 0; r < reps + 2; ++r) {
            CK(cudaEventRecord(e0));
            gather_ldg_kernel<8><<<grid, 256>>>(H, out, c.n, c.ld, lanes, c.deg, log2_dmax, thr, 1234u + r,
                                                c.rows_per_cta, c.write_out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float t;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (r >= 2) ms.push_back(t);
        }
        std::sort(ms.begin(), ms.end());
        const double med = ms[ms.size() / 2];
        const double gathers = (double)c.n * c.deg;
        const double gbytes = gathers * c.F * 4.0;
        printf("{\"bench\": \"gather_ldg\", \"name\": \"%s\", \"n\": %d, \"F\": %d, \"ld\": %d, \"deg\": %d, "
               "\"dmax\": %.0f, \"p_uniform\": %.2f, \"write_out\": %d, \"rows_per_cta\": %d, \"ms\": %.4f, \"ms_min\": %.4f, "
               "\"gather_rows_per_s\": %.4g, \"gather_GBps\": %.1f, \"H_MB\": %.1f}\n",
               c.name, c.n, c.F, c.ld, c.deg, dmax, c.p_uni, c.write_out, c.rows_per_cta, med, ms[0],
               gathers / (med * 1e-3), gbytes / (med * 1e-3) / 1e9, (double)c.n * c.ld * 4 / 1e6);
        fflush(stdout);
    }
    return 0;
}
