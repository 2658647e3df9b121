// CSR SpMM for sm_100a: row-split main kernel + split long rows (nnz-chunked, fixed-order reduce).
//
// Replaces tf.sparse.sparse_dense_matmul at gnntf/core/gnn/architectures/filter.py:19 and
// gcn.py:24,48,88,104,131, fused with the teleport axpy / dropout / activation of
// filter.py:21-22.  The path is bandwidth-bound (AI < 7 flop/B): no tensor cores.
//
// Lane mapping.  A feature row of F floats is cut into VEC-wide slots (VEC = 4 -> LDG.128 when
// every leading dimension and base pointer allows it, else scalar).  GROUP lanes (a power of two
// <= 32) own one sparse row; lane `gl` of the group owns slots gl, gl+GROUP, ... (NSLOT of them),
// so the lanes of a group read one contiguous GROUP*VEC*4-byte span of each gathered row.  A warp
// therefore runs 32/GROUP sparse rows at once.  The group's lanes fetch GROUP (col,val) pairs with
// one coalesced streaming load each and broadcast them through a small per-warp shared-memory
// buffer; the gathers of UNROLL consecutive entries are issued back to back before the FMAs
// (memory-level parallelism).  Rows wider than 128 floats run as 128-float tiles over grid.y.
//
// Accumulation order inside a row is CSR slot order = COO storage order (the builder's sort is
// stable), i.e. the order TF's CPU kernel uses.  Rows longer than A->long_threshold are skipped
// by the row warps and run as fixed-size pieces on other warps of the same grid (process_piece),
// whose partial sums are combined in piece order by spmm_long_reduce_kernel: deterministic, no
// float atomics.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "spmm.cuh"

namespace gnntf {

template <int VEC>
__device__ __forceinline__ void apply_epilogue(const Epilogue& e, int64_t row, int f, Vec<VEC> acc) {
    Vec<VEC> out;
#pragma unroll
    for (int i = 0; i < VEC; ++i) out.v[i] = acc.v[i] * e.s;
    if (e.H0 != nullptr) {
        Vec<VEC> h = Vec<VEC>::stream(e.H0 + row * e.ldh + f);
#pragma unroll
        for (int i = 0; i < VEC; ++i) out.v[i] = __fadd_rn(out.v[i], __fmul_rn(h.v[i], e.t));
    }
    if (e.keep != nullptr) {
        const uint8_t* k = e.keep + row * (int64_t)e.F + f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) out.v[i] = k[i] ? out.v[i] * e.p_scale : 0.0f;
    }
    if (e.act == GNNTF_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) out.v[i] = fmaxf(out.v[i], 0.0f);
    }
    if (e.C != nullptr) out.store(e.C + row * e.ldc + f);
    if (e.ACC != nullptr) {
        float* a = e.ACC + row * e.ldacc + f;
        Vec<VEC> b;
#pragma unroll
        for (int i = 0; i < VEC; ++i) b.v[i] = 0.0f;
        if (e.u != 0.0f) b = Vec<VEC>::stream(e.B + row * e.ldb + f);  // the dense operand's own row
        Vec<VEC> r;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float prev = e.acc_init ? 0.0f : a[i];
            r.v[i] = prev + e.u * b.v[i] + e.w * out.v[i];
        }
        r.store(a);
    }
}

// L2 eviction-policy helpers: data touched once per step (CSR arrays, the teleport term, the
// output) is marked evict-first so it does not displace gathered feature rows from L2.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int ld_once(const int* p, uint64_t pol) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_once(const float* p, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ld_once4(const float* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_once4(float* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
}


// ---------------------------------------------------------------------------------------------
// Long rows, phase 1: one warp per 256-entry piece.  The 32/GROUP lane groups take the piece's
// entries round-robin and are combined with xor-shuffles in a fixed pattern; the result goes to
// partials[piece, ldp] (ldp = round_up(F,4)).  Runs inside the row kernel's grid (the first
// `piece_ctas` CTAs), so the pieces overlap the ordinary rows instead of being a second launch.
// ---------------------------------------------------------------------------------------------
struct PieceArgs {
    const int* chunk_row;
    const int* chunk_begin;
    int n_chunks;
    int chunk;
    float* partials;
    int ldp;
};

template <int VEC, int NSLOT, int GROUP, int UNROLL>
__device__ __forceinline__ void process_piece(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                              const float* __restrict__ val, const float* __restrict__ B,
                                              int64_t ldb, const int* __restrict__ chunk_row,
                                              const int* __restrict__ chunk_begin, int piece, int chunk,
                                              float* __restrict__ partials, int ldp, int F, int f_base) {
    constexpr int NGROUPS = 32 / GROUP;
    const int lane = threadIdx.x & 31;
    const int g = lane / GROUP;
    const int gl = lane % GROUP;
    const int row = __ldg(chunk_row + piece);
    const int begin = __ldg(chunk_begin + piece);
    const int end = min(begin + chunk, __ldg(row_ptr + row + 1));

    int fo[NSLOT];
    bool fok[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        fo[s] = f_base + (s * GROUP + gl) * VEC;
        fok[s] = fo[s] < F;
    }
    Vec<VEC> acc[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[s].v[i] = 0.0f;

    // 32 entries per outer step: lane l holds entry base+l; group g consumes entries with
    // (index % NGROUPS) == g.
    for (int base = begin; base < end; base += 32) {
        int c = 0;
        float v = 0.0f;
        if (base + lane < end) {
            c = ld_stream(col_idx + base + lane);
            v = ld_stream(val + base + lane);
        }
        const int cnt = min(32, end - base);
        for (int j = 0; j < cnt; j += NGROUPS * UNROLL) {
            int cj[UNROLL];
            float vj[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int src = j + u * NGROUPS + g;
                cj[u] = __shfl_sync(0xffffffffu, c, src & 31);
                vj[u] = __shfl_sync(0xffffffffu, v, src & 31);
                if (src >= cnt) vj[u] = 0.0f, cj[u] = -1;
            }
            Vec<VEC> x[UNROLL][NSLOT];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const float* src = B + (int64_t)cj[u] * ldb;
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) {
                    if (cj[u] >= 0 && fok[s]) {
                        x[u][s] = Vec<VEC>::gather(src + fo[s]);
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) x[u][s].v[i] = 0.0f;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[s].v[i] = fmaf(vj[u], x[u][s].v[i], acc[s].v[i]);
        }
    }
    if (NGROUPS > 1) {
#pragma unroll
        for (int o = GROUP; o < 32; o <<= 1)
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[s].v[i] += __shfl_xor_sync(0xffffffffu, acc[s].v[i], o);
    }
    if (g == 0) {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
            if (fok[s]) acc[s].store(partials + (int64_t)piece * ldp + fo[s]);
    }
}

// ---------------------------------------------------------------------------------------------
// Main kernel: GROUP lanes per sparse row; rows with deg > long_threshold are left to the piece
// path.  A CTA walks blocks_per_cta consecutive row blocks (fewer, longer-lived CTAs: the per-row
// version launched 306 k CTAs on the products shape and ran at 47 % achieved occupancy);
// grid.y walks feature tiles of GROUP*NSLOT*VEC floats.
// A gather address costs one mad.wide.u32 (the first profile showed 28 executed instructions per
// entry, most of them 64-bit address arithmetic, per-entry predicates and register zeroing); every
// batch of UNROLL entries runs predicate-free (short rows are padded, see the loop).
// ---------------------------------------------------------------------------------------------

// Address of a gathered row piece in ONE instruction: column ids are non-negative int32 and the
// row pitch in bytes fits 32 bits, so base + c*pitch is exactly mad.wide.u32 (IMAD.WIDE.U32).
template <int VEC>
__device__ __forceinline__ Vec<VEC> gather_row(const float* lane_base, int c, uint32_t pitch_bytes) {
    const float* p;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(c), "r"(pitch_bytes), "l"(lane_base));
    return Vec<VEC>::gather(p);
}

template <int VEC, int NSLOT, int GROUP, int UNROLL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
spmm_rows_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                 const float* __restrict__ val, const int* __restrict__ row_map,
                 const float* __restrict__ B, int64_t ldb, int n_rows, int long_threshold,
                 int blocks_per_cta, int piece_ctas, PieceArgs pieces, Epilogue epi) {
    static_assert(UNROLL % 2 == 0 && GROUP % UNROLL == 0, "entries are read back two at a time, whole batches");
    if ((int)blockIdx.x < piece_ctas) {  // CTA-uniform: this CTA works on pieces of split rows
        const int piece = blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
        if (piece < pieces.n_chunks)
            process_piece<VEC, NSLOT, GROUP, (NSLOT >= 4) ? 2 : 4>(row_ptr, col_idx, val, B, ldb, pieces.chunk_row,
                                                                  pieces.chunk_begin, piece, pieces.chunk,
                                                                  pieces.partials, pieces.ldp, epi.F,
                                                                  blockIdx.y * (GROUP * NSLOT * VEC));
        return;
    }
    const int row_cta = blockIdx.x - piece_ctas;
    constexpr int ROWS_PER_WARP = 32 / GROUP;
    constexpr int WARPS = THREADS / 32;
    constexpr int ROWS_PER_BLOCK = WARPS * ROWS_PER_WARP;
    // (col,val) pairs of the current 32 entries of this warp, broadcast through shared memory:
    // one LDS.128 delivers two entries to every lane (0.5 L1 wavefronts per entry; the shuffle
    // pair it replaces cost 2 wavefronts and was a third of the L1 data-pipe traffic, profiles/r1/02)
    __shared__ __align__(16) int2 cv_smem[WARPS][32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int g = lane / GROUP;
    const int gl = lane % GROUP;
    const int f_base = blockIdx.y * (GROUP * NSLOT * VEC);
    const int F = epi.F;
    const uint64_t pol = policy_evict_first();
    const uint32_t pitch = (uint32_t)ldb * 4u;
    int2* cv = cv_smem[warp];
    const int2* cv_group = cv + g * GROUP;

    int fo[NSLOT];
    bool fok[NSLOT];
    const float* Bl[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        fo[s] = f_base + (s * GROUP + gl) * VEC;
        fok[s] = fo[s] < F;  // VEC == 4 implies F % 4 == 0, so the whole slot is in range
        Bl[s] = B + (fok[s] ? fo[s] : 0);
    }

    for (int it = 0; it < blocks_per_cta; ++it) {
        const int64_t row = ((int64_t)row_cta * blocks_per_cta + it) * ROWS_PER_BLOCK + warp * ROWS_PER_WARP + g;
        if (row - g >= n_rows) break;  // warp-uniform: the whole warp is past the end
        int start = 0, deg = 0;
        bool mine = false;
        if (row < n_rows) {
            start = __ldg(row_ptr + row);
            deg = __ldg(row_ptr + row + 1) - start;
            mine = !(long_threshold > 0 && deg > long_threshold);
            if (!mine) deg = 0;
        }
        int maxdeg = deg;
        if (ROWS_PER_WARP > 1) {
#pragma unroll
            for (int o = GROUP; o < 32; o <<= 1) {
                maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, o));
            }
        }
        Vec<VEC> acc[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[s].v[i] = 0.0f;

        // Slots past the end of a row are padded with (first column of the row, value 0): the
        // batches then run predicate-free for every group of the warp, whatever its degree; the
        // padded gathers re-read a row that is already in L1.  (0 * x leaves the sum unchanged for
        // finite x; a non-finite feature row of the first neighbour already makes the true result
        // non-finite.)
        int c_pad = 0;
        for (int off = 0; off < maxdeg; off += GROUP) {
            int c = c_pad;
            float v = 0.0f;
            if (off + gl < deg) {
                c = ld_once(col_idx + start + off + gl, pol);
                v = ld_once(val + start + off + gl, pol);
            }
            if (off == 0) {
                c_pad = __shfl_sync(0xffffffffu, c, 0, GROUP);  // the row's first column (deg > 0)
                if (gl >= deg) c = (deg > 0) ? c_pad : 0;
            }
            __syncwarp();  // everyone is done reading the previous batch
            cv[lane] = make_int2(c, __float_as_int(v));
            __syncwarp();
            const int lim = min(GROUP, maxdeg - off);
            for (int j = 0; j < lim; j += UNROLL) {
                int cj[UNROLL];
                float vj[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u += 2) {
                    const int4 e = *reinterpret_cast<const int4*>(cv_group + j + u);
                    cj[u] = e.x; vj[u] = __int_as_float(e.y);
                    cj[u + 1] = e.z; vj[u + 1] = __int_as_float(e.w);
                }
                Vec<VEC> x[UNROLL][NSLOT];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s)
                        if (fok[s]) x[u][s] = gather_row<VEC>(Bl[s], cj[u], pitch);
#pragma unroll
                for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s)
                        if (fok[s])
#pragma unroll
                            for (int i = 0; i < VEC; ++i) acc[s].v[i] = fmaf(vj[u], x[u][s].v[i], acc[s].v[i]);
            }
        }
        if (deg == 0) {  // an empty (or split) row riding along in a warp: drop whatever the padding produced
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[s].v[i] = 0.0f;
        }
        if (mine) {
            const int64_t out_row = row_map ? (int64_t)__ldg(row_map + row) : row;
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
                if (fok[s]) apply_epilogue<VEC>(epi, out_row, fo[s], acc[s]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Bulk-copy kernel (wide rows: F >= 68 floats, 16-byte aligned rows).
//
// Why it exists (profiles/r1/01): on the products shape the per-row register
// kernel is latency-bound — DRAM, L1 and the issue slots are each ~50 % busy — because the bytes a
// warp can keep in flight are capped by registers and by the 6 counting scoreboards (a rolling
// register ring was measured 2x SLOWER: waiting for the oldest load also waits for the youngest;
// a shared-memory ring fed by 16-byte cp.async was 2.5x slower: LDGSTS.128 sustains ~16 B/clk/SM).
// Here every gathered feature row travels global -> shared as ONE bulk async copy (PTX
// cp.async.bulk, SASS UBLKCP: the TMA engine, no registers, no L1, no scoreboard) that signals an
// mbarrier with its byte count.  Each warp owns a private ring of S stages x G row images and S
// mbarriers; it keeps up to S*G gathers in flight across row boundaries, so bytes in flight are
// bounded by shared memory (~200 KB/SM), not by the register file.
//
// A warp owns R consecutive rows at a time and walks such chunks with a grid stride: the grid
// sweeps the matrix as one wavefront, which keeps the band of gathered rows that L2 must hold
// narrow.  Rows above the split threshold are left to the piece kernels, as in spmm_rows_kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// A byte-count mismatch would spin forever and wedge the GPU: trap after ~2 s instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// Epilogue with the teleport row already in registers and evict-first stores; the rarely used
// dropout-mask / backward-accumulate variants go through the generic path.
__device__ __forceinline__ void epilogue_fast(const Epilogue& e, int64_t row, int f, const float4& a,
                                              const float4& h0, uint64_t pol) {
    if (e.keep != nullptr || e.ACC != nullptr) {
        apply_epilogue<4>(e, row, f, Vec<4>{{a.x, a.y, a.z, a.w}});
        return;
    }
    float4 o = make_float4(a.x * e.s, a.y * e.s, a.z * e.s, a.w * e.s);
    if (e.H0 != nullptr) {
        o.x = __fadd_rn(o.x, __fmul_rn(h0.x, e.t));
        o.y = __fadd_rn(o.y, __fmul_rn(h0.y, e.t));
        o.z = __fadd_rn(o.z, __fmul_rn(h0.z, e.t));
        o.w = __fadd_rn(o.w, __fmul_rn(h0.w, e.t));
    }
    if (e.act == GNNTF_ACT_RELU) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    if (e.C != nullptr) st_once4(e.C + row * e.ldc + f, o, pol);
}

// G row images per stage, S stages per warp.  Entry e of a stream lives in stage (e/G)%S, slot e%G.
template <int NSLOT, int G, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
spmm_bulk_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                 const float* __restrict__ val, const int* __restrict__ row_map,
                 const float* __restrict__ B, int64_t ldb, int n_rows, int long_threshold,
                 int rows_per_chunk, int tile_floats, Epilogue epi) {
    static_assert(32 % G == 0 && G <= 32, "G must divide 32");
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int F = epi.F;
    const int f_base = blockIdx.y * tile_floats;
    const int width = min(tile_floats, F - f_base);        // floats of this feature tile
    const uint32_t row_bytes = (uint32_t)width * 4u;       // multiple of 16 (F % 4 == 0)
    const uint32_t slot_bytes = (uint32_t)tile_floats * 4u;
    const uint64_t pol = policy_evict_first();

    unsigned char* my_ring = bulk_smem + (size_t)warp * (S * G) * slot_bytes;
    const uint32_t ring_u32 = smem_addr(my_ring);
    const uint32_t bars_u32 = smem_addr(bulk_smem + (size_t)WARPS * (S * G) * slot_bytes) + warp * S * 8;
    if (lane < S) mbar_init(bars_u32 + lane * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    int fo[NSLOT];
    bool fok[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int in_tile = (s * 32 + lane) * 4;
        fo[s] = f_base + in_tile;
        fok[s] = in_tile < width;
    }
    const float* Bt = B + f_base;
    uint32_t phase_bits = 0;  // bit s = parity to wait for on stage s next

    const int n_chunks = (n_rows + rows_per_chunk - 1) / rows_per_chunk;
    const int warp_stride = gridDim.x * WARPS;
    float4 acc[NSLOT], h0[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) h0[s] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int chunk = blockIdx.x * WARPS + warp; chunk < n_chunks; chunk += warp_stride) {
        const int r0 = chunk * rows_per_chunk;
        const int nr = min(rows_per_chunk, n_rows - r0);
        const int rp = (lane <= nr) ? __ldg(row_ptr + r0 + lane) : 0;  // lane i holds row_ptr[r0+i]
        const int rp_next = __shfl_down_sync(0xffffffffu, rp, 1);
        const bool is_long = (lane < nr) && long_threshold > 0 && (rp_next - rp) > long_threshold;
        const unsigned long_mask = __ballot_sync(0xffffffffu, is_long);
        int cur_row = 0, cur_end = 0;

        auto prefetch_h0 = [&](int local_row) {
            if (epi.H0 != nullptr) {
                const int64_t m = row_map ? (int64_t)__ldg(row_map + r0 + local_row) : (int64_t)(r0 + local_row);
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
                    if (fok[s]) h0[s] = ld_once4(epi.H0 + m * epi.ldh + fo[s], pol);
            }
        };
        auto flush_row = [&](int seg_end) {  // write the current row, advance, prefetch the next teleport row
            const int64_t m = row_map ? (int64_t)__ldg(row_map + r0 + cur_row) : (int64_t)(r0 + cur_row);
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) {
                if (fok[s]) epilogue_fast(epi, m, fo[s], acc[s], h0[s], pol);
                acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            ++cur_row;
            cur_end = __shfl_sync(0xffffffffu, rp, min(cur_row + 1, 31));
            if (cur_row < seg_end) prefetch_h0(cur_row);
        };

        int a = 0;
        while (a < nr) {  // maximal runs [a,b) of rows that are not split
            if ((long_mask >> a) & 1u) { ++a; continue; }
            int b = nr;
            const unsigned rest = long_mask >> a;
            if (rest) b = a + __ffs(rest) - 1;
            const int P0 = __shfl_sync(0xffffffffu, rp, a);
            const int P1 = __shfl_sync(0xffffffffu, rp, b);
            cur_row = a;
            cur_end = __shfl_sync(0xffffffffu, rp, a + 1);
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
            prefetch_h0(a);
            while (cur_row < b && cur_end == P0) flush_row(b);  // leading empty rows
            if (P0 == P1) { a = b; continue; }

            const int n_ent = P1 - P0;
            const int n_groups = (n_ent + G - 1) / G;
            // (col,val) batches of 32 entries; *_nx prefetched one batch ahead of their first use
            int cb = 0, cb_nx = 0;     // issue side
            float vb = 0.f, vb_nx = 0.f;  // consume side
            if (lane < n_ent) { cb = ld_once(col_idx + P0 + lane, pol); vb = ld_once(val + P0 + lane, pol); }
            if (32 + lane < n_ent) { cb_nx = ld_once(col_idx + P0 + 32 + lane, pol); vb_nx = ld_once(val + P0 + 32 + lane, pol); }
            int cb_base = 0;  // stream index of cb's lane 0

            auto issue_group = [&](int g) {  // copies of group g into stage g % S
                const int e0 = g * G;
                if (e0 >= cb_base + 32) {  // the issue side moves into the next batch
                    cb = cb_nx;
                    cb_base += 32;
                    cb_nx = (cb_base + 32 + lane < n_ent) ? ld_once(col_idx + P0 + cb_base + 32 + lane, pol) : 0;
                }
                const int cnt = min(G, n_ent - e0);
                const int stage = g % S;
                const uint32_t bar = bars_u32 + stage * 8;
                if (lane == 0) mbar_expect_tx(bar, (uint32_t)cnt * row_bytes);
                __syncwarp();
                const int c = __shfl_sync(0xffffffffu, cb, (e0 - cb_base + lane) & 31);
                if (lane < cnt)
                    bulk_g2s(ring_u32 + (uint32_t)(stage * G + lane) * slot_bytes, Bt + (int64_t)c * ldb, row_bytes, bar);
            };

            const int pre = min(S, n_groups);
            for (int g = 0; g < pre; ++g) issue_group(g);
            int vb_base = 0;
            for (int g = 0; g < n_groups; ++g) {
                const int stage = g % S;
                const int e0 = g * G;
                if (e0 >= vb_base + 32) {
                    vb = vb_nx;
                    vb_base += 32;
                    vb_nx = (vb_base + 32 + lane < n_ent) ? ld_once(val + P0 + vb_base + 32 + lane, pol) : 0.f;
                }
                mbar_wait(bars_u32 + stage * 8, (phase_bits >> stage) & 1u);
                phase_bits ^= (1u << stage);
                const int cnt = min(G, n_ent - e0);
                const unsigned char* stage_base = my_ring + (size_t)(stage * G) * slot_bytes;
                if (cnt == G && cur_end > P0 + e0 + G) {  // full group, no row end inside: check-free
#pragma unroll
                    for (int k = 0; k < G; ++k) {
                        const float v = __shfl_sync(0xffffffffu, vb, (e0 - vb_base + k) & 31);
#pragma unroll
                        for (int s = 0; s < NSLOT; ++s) {
                            if (fok[s]) {
                                const float4 x = *reinterpret_cast<const float4*>(stage_base + (size_t)k * slot_bytes + (s * 32 + lane) * 16);
                                acc[s].x = fmaf(v, x.x, acc[s].x);
                                acc[s].y = fmaf(v, x.y, acc[s].y);
                                acc[s].z = fmaf(v, x.z, acc[s].z);
                                acc[s].w = fmaf(v, x.w, acc[s].w);
                            }
                        }
                    }
                } else {
                    for (int k = 0; k < cnt; ++k) {
                        const float v = __shfl_sync(0xffffffffu, vb, (e0 - vb_base + k) & 31);
#pragma unroll
                        for (int s = 0; s < NSLOT; ++s) {
                            if (fok[s]) {
                                const float4 x = *reinterpret_cast<const float4*>(stage_base + (size_t)k * slot_bytes + (s * 32 + lane) * 16);
                                acc[s].x = fmaf(v, x.x, acc[s].x);
                                acc[s].y = fmaf(v, x.y, acc[s].y);
                                acc[s].z = fmaf(v, x.z, acc[s].z);
                                acc[s].w = fmaf(v, x.w, acc[s].w);
                            }
                        }
                        while (cur_row < b && cur_end == P0 + e0 + k + 1) flush_row(b);
                    }
                }
                __syncwarp();  // every lane is done reading this stage before it is refilled
                if (g + S < n_groups) issue_group(g + S);
            }
            a = b;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Stand-alone launch of the piece work (used by the bulk-copy variant; the row kernel runs the
// pieces inside its own grid).
template <int VEC, int NSLOT, int GROUP, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS)
spmm_chunk_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                  const float* __restrict__ val, const float* __restrict__ B, int64_t ldb,
                  const int* __restrict__ chunk_row, const int* __restrict__ chunk_begin,
                  int n_chunks, int chunk, float* __restrict__ partials, int ldp, int F) {
    const int piece = blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    if (piece >= n_chunks) return;  // warp-uniform
    process_piece<VEC, NSLOT, GROUP, UNROLL>(row_ptr, col_idx, val, B, ldb, chunk_row, chunk_begin, piece, chunk,
                                             partials, ldp, F, blockIdx.y * (GROUP * NSLOT * VEC));
}

// Long rows, phase 2: one CTA per long row; thread t owns feature t (strided), sums the row's
// pieces in piece order, then applies the epilogue.
__global__ void __launch_bounds__(128)
spmm_long_reduce_kernel(const int* __restrict__ long_row, const int* __restrict__ long_first_chunk,
                        const int* __restrict__ long_n_chunks, const int* __restrict__ row_map,
                        const float* __restrict__ partials, int ldp, Epilogue epi) {
    const int li = blockIdx.x;
    const int csr_row = __ldg(long_row + li);
    const int64_t row = row_map ? (int64_t)__ldg(row_map + csr_row) : (int64_t)csr_row;
    const int first = __ldg(long_first_chunk + li);
    const int cnt = __ldg(long_n_chunks + li);
    for (int f = threadIdx.x; f < epi.F; f += blockDim.x) {
        const float* p = partials + (int64_t)first * ldp + f;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // 4 independent chains, fixed assignment
        int k = 0;
        for (; k + 4 <= cnt; k += 4) {
            a0 += ld_stream(p + (int64_t)(k + 0) * ldp);
            a1 += ld_stream(p + (int64_t)(k + 1) * ldp);
            a2 += ld_stream(p + (int64_t)(k + 2) * ldp);
            a3 += ld_stream(p + (int64_t)(k + 3) * ldp);
        }
        for (; k < cnt; ++k) a0 += ld_stream(p + (int64_t)k * ldp);
        Vec<1> acc{{(a0 + a1) + (a2 + a3)}};
        apply_epilogue<1>(epi, row, f, acc);
    }
}

// ---------------------------------------------------------------------------------------------
// Host-side dispatch
// ---------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
// Tuning knobs for A/B measurements (read once): GNNTF_SPMM_BULK=1 selects the TMA bulk-copy
// kernel for wide rows (measured SLOWER than the register kernel on B200: 102 vs 61 ms on the
// products shape — one 400-byte bulk copy costs ~23 cycles of TMA issue per SM — so it is off by
// default), GNNTF_SPMM_ROWS the rows per chunk.
static int rows_blocks_per_cta() { static int v = std::max(1, env_int("GNNTF_SPMM_BPC", 8)); return v; }
static int bulk_mode() { static int v = env_int("GNNTF_SPMM_BULK", 0); return v; }
static int bulk_rows_per_chunk() { static int v = std::max(1, std::min(31, env_int("GNNTF_SPMM_ROWS", 16))); return v; }

template <int VEC, int NSLOT, int GROUP>
static int launch_cfg(const gnntf_csr_t* A, const float* B, int64_t ldb, const Epilogue& epi,
                      cudaStream_t st) {
    constexpr int THREADS = 256;
    constexpr int UNROLL = (NSLOT >= 4) ? 2 : (NSLOT == 2 ? 4 : (GROUP >= 8 ? 8 : 4));
    constexpr int ROWS_PER_CTA = (THREADS / 32) * (32 / GROUP);
    const int F = epi.F;
    const int tile = GROUP * NSLOT * VEC;
    const unsigned gy = (unsigned)ceil_div(F, tile);
    const int thr = (A->n_long > 0) ? A->long_threshold : 0;
    if (A->n_rows > 0) {
        // several row blocks per CTA once the grid is deep enough to keep every SM busy for many
        // waves; small graphs (Cora: 43 row blocks) keep one block per CTA
        const int64_t row_blocks = ceil_div(A->n_rows, ROWS_PER_CTA);
        const int bpc = (int)std::max<int64_t>(1, std::min<int64_t>(rows_blocks_per_cta(), row_blocks / ((int64_t)kNumSMs * 5 * 4)));
        // the pieces of split rows ride in the same grid, ahead of the ordinary rows
        PieceArgs pieces{};
        int piece_ctas = 0;
        const int ldp = (int)round_up(F, 4);
        if (A->n_long > 0) {
            pieces = PieceArgs{A->chunk_row, A->chunk_begin, A->n_chunks, A->chunk, A->partials, ldp};
            piece_ctas = (int)ceil_div(A->n_chunks, THREADS / 32);
        }
        dim3 grid((unsigned)(piece_ctas + ceil_div(A->n_rows, (int64_t)ROWS_PER_CTA * bpc)), gy);
        spmm_rows_kernel<VEC, NSLOT, GROUP, UNROLL, THREADS, 5><<<grid, THREADS, 0, st>>>(
            A->row_ptr, A->col_idx, A->val, A->row_map, B, ldb, (int)A->n_rows, thr, bpc, piece_ctas, pieces, epi);
        GNNTF_LAUNCH_CHECK();
        if (A->n_long > 0) {
            spmm_long_reduce_kernel<<<A->n_long, 128, 0, st>>>(A->long_row, A->long_first_chunk,
                                                              A->long_n_chunks, A->row_map, A->partials,
                                                              ldp, epi);
            GNNTF_LAUNCH_CHECK();
        }
    }
    return GNNTF_OK;
}


template <int NSLOT, int G, int S>
static int launch_bulk(const gnntf_csr_t* A, const float* B, int64_t ldb, const Epilogue& epi,
                       cudaStream_t st) {
    const int F = epi.F;
    const int tile_floats = std::min(F, NSLOT * 128);
    const unsigned gy = (unsigned)ceil_div(F, tile_floats);
    const size_t per_warp = (size_t)S * G * tile_floats * 4;
    const int thr = (A->n_long > 0) ? A->long_threshold : 0;
    const int rows_per_chunk = bulk_rows_per_chunk();
    auto go = [&](auto warps_tag) -> int {
        constexpr int WARPS = decltype(warps_tag)::value;
        const size_t smem = per_warp * WARPS + (size_t)WARPS * S * 8;
        auto kern = spmm_bulk_kernel<NSLOT, G, S, WARPS>;
        GNNTF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t chunks = ceil_div(A->n_rows, rows_per_chunk);
        const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(224 * 1024) / (smem + 1024)));
        dim3 grid((unsigned)std::min<int64_t>(ceil_div(chunks, WARPS), (int64_t)kNumSMs * ctas_per_sm), gy);
        kern<<<grid, WARPS * 32, smem, st>>>(A->row_ptr, A->col_idx, A->val, A->row_map, B, ldb,
                                            (int)A->n_rows, thr, rows_per_chunk, tile_floats, epi);
        GNNTF_LAUNCH_CHECK();
        return GNNTF_OK;
    };
    if (A->n_rows > 0) {
        int rc;
        if (per_warp * 8 <= 110 * 1024) rc = go(std::integral_constant<int, 8>{});
        else if (per_warp * 4 <= 110 * 1024) rc = go(std::integral_constant<int, 4>{});
        else rc = go(std::integral_constant<int, 2>{});
        if (rc != GNNTF_OK) return rc;
    }
    if (A->n_long > 0) {
        constexpr int THREADS = 256;
        const int ldp = (int)round_up(F, 4);
        constexpr int RS = (NSLOT > 4) ? 4 : NSLOT;
        const unsigned gy2 = (unsigned)ceil_div(F, 32 * RS * 4);
        dim3 grid((unsigned)ceil_div(A->n_chunks, THREADS / 32), gy2);
        spmm_chunk_kernel<4, RS, 32, (RS >= 4) ? 2 : 4, THREADS><<<grid, THREADS, 0, st>>>(
            A->row_ptr, A->col_idx, A->val, B, ldb, A->chunk_row, A->chunk_begin, A->n_chunks,
            A->chunk, A->partials, ldp, F);
        GNNTF_LAUNCH_CHECK();
        spmm_long_reduce_kernel<<<A->n_long, 128, 0, st>>>(A->long_row, A->long_first_chunk,
                                                          A->long_n_chunks, A->row_map, A->partials,
                                                          ldp, epi);
        GNNTF_LAUNCH_CHECK();
    }
    return GNNTF_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int validate_csr(const gnntf_csr_t* A) {
    if (A == nullptr) return GNNTF_E_NULL;
    if (A->n_rows < 0 || A->nnz < 0 || A->nnz > 0x7fffffffLL || A->n_rows > 0x7ffffffeLL) return GNNTF_E_SIZE;
    if (A->n_rows > 0 && A->row_ptr == nullptr) return GNNTF_E_NULL;
    if (A->nnz > 0 && (A->col_idx == nullptr || A->val == nullptr)) return GNNTF_E_NULL;
    if (A->n_long < 0 || A->n_chunks < 0) return GNNTF_E_SIZE;
    if (A->n_long > 0) {
        if (A->long_row == nullptr || A->long_first_chunk == nullptr || A->long_n_chunks == nullptr ||
            A->chunk_row == nullptr || A->chunk_begin == nullptr || A->partials == nullptr)
            return GNNTF_E_NULL;
        if (A->long_threshold <= 0 || A->chunk <= 0) return GNNTF_E_SIZE;
    }
    return GNNTF_OK;
}

// C (and ACC) = epilogue(A·B).  Chooses the lane mapping from F and the alignment of every
// operand; see the file header.
int spmm_dispatch(const gnntf_csr_t* A, const float* B, int64_t ldb, Epilogue epi, cudaStream_t st) {
    int rc = validate_csr(A);
    if (rc != GNNTF_OK) return rc;
    const int64_t F = epi.F;
    if (F < 0 || F > 0x7fffffff) return GNNTF_E_SIZE;
    if (A->n_rows == 0 || F == 0) return GNNTF_OK;
    if (B == nullptr) return GNNTF_E_NULL;
    if (ldb >= (1LL << 30)) return GNNTF_E_SIZE;  // row pitch in bytes must fit 32 bits
    if (ldb < F || (epi.C && epi.ldc < F) || (epi.H0 && epi.ldh < F) || (epi.ACC && epi.ldacc < F))
        return GNNTF_E_SIZE;
    epi.B = B;
    epi.ldb = ldb;

    bool v4 = (F % 4 == 0) && (ldb % 4 == 0) && aligned16(B);
    if (epi.C) v4 = v4 && (epi.ldc % 4 == 0) && aligned16(epi.C);
    if (epi.H0) v4 = v4 && (epi.ldh % 4 == 0) && aligned16(epi.H0);
    if (epi.ACC) v4 = v4 && (epi.ldacc % 4 == 0) && aligned16(epi.ACC);
    if (epi.keep) v4 = false;  // byte mask rows are F-strided: keep the scalar path

    if (v4 && bulk_mode() && F >= 68) {
        if (F <= 128) return launch_bulk<1, 8, 4>(A, B, ldb, epi, st);
        if (F <= 256) return launch_bulk<2, 8, 2>(A, B, ldb, epi, st);
        return launch_bulk<4, 4, 2>(A, B, ldb, epi, st);  // tiles of 512 floats over grid.y
    }
    if (v4) {
        const int64_t slots = F / 4;
        if (slots <= 4) return launch_cfg<4, 1, 4>(A, B, ldb, epi, st);
        if (slots <= 8) return launch_cfg<4, 1, 8>(A, B, ldb, epi, st);
        if (slots <= 16) return launch_cfg<4, 1, 16>(A, B, ldb, epi, st);
        if (slots <= 32) return launch_cfg<4, 1, 32>(A, B, ldb, epi, st);
        // wider rows: 128-float tiles over grid.y with the one-slot mapping (8-deep gather batches);
        // GNNTF_SPMM_WIDE=0 selects the multi-slot mappings instead (A/B)
        static const int wide_tiles = env_int("GNNTF_SPMM_WIDE", 1);
        if (wide_tiles) return launch_cfg<4, 1, 32>(A, B, ldb, epi, st);
        if (slots <= 64) return launch_cfg<4, 2, 32>(A, B, ldb, epi, st);
        return launch_cfg<4, 4, 32>(A, B, ldb, epi, st);  // tiles of 512 floats over grid.y
    }
    if (F <= 4) return launch_cfg<1, 1, 4>(A, B, ldb, epi, st);
    if (F <= 8) return launch_cfg<1, 1, 8>(A, B, ldb, epi, st);
    if (F <= 16) return launch_cfg<1, 1, 16>(A, B, ldb, epi, st);
    if (F <= 32) return launch_cfg<1, 1, 32>(A, B, ldb, epi, st);
    if (F <= 64) return launch_cfg<1, 2, 32>(A, B, ldb, epi, st);
    return launch_cfg<1, 4, 32>(A, B, ldb, epi, st);  // tiles of 128 floats over grid.y
}

// ---------------------------------------------------------------------------------------------
// Long-row plan
// ---------------------------------------------------------------------------------------------
__global__ void plan_count_kernel(const int* __restrict__ row_ptr, int n_rows, int thr, int chunk,
                                  int* __restrict__ counts) {
    int nl = 0, nc = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int deg = row_ptr[r + 1] - row_ptr[r];
        if (deg > thr) {
            nl += 1;
            nc += (deg + chunk - 1) / chunk;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nl += __shfl_xor_sync(0xffffffffu, nl, o);
        nc += __shfl_xor_sync(0xffffffffu, nc, o);
    }
    if ((threadIdx.x & 31) == 0 && nl > 0) {
        atomicAdd(counts + 0, nl);
        atomicAdd(counts + 1, nc);
    }
}

// Long rows are rare; one thread claims a contiguous range of pieces for its row.  The order in
// which rows claim ranges is arbitrary but never changes a result (pieces of one row stay
// contiguous and ordered).
__global__ void plan_fill_kernel(const int* __restrict__ row_ptr, int n_rows, int thr, int chunk,
                                 int* __restrict__ counters, int* __restrict__ long_row,
                                 int* __restrict__ long_first_chunk, int* __restrict__ long_n_chunks,
                                 int* __restrict__ chunk_row, int* __restrict__ chunk_begin) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int start = row_ptr[r];
        const int deg = row_ptr[r + 1] - start;
        if (deg > thr) {
            const int nch = (deg + chunk - 1) / chunk;
            const int li = atomicAdd(counters + 0, 1);
            const int first = atomicAdd(counters + 1, nch);
            long_row[li] = (int)r;
            long_first_chunk[li] = first;
            long_n_chunks[li] = nch;
            for (int j = 0; j < nch; ++j) {
                chunk_row[first + j] = (int)r;
                chunk_begin[first + j] = start + j * chunk;
            }
        }
    }
}

}  // namespace gnntf

using namespace gnntf;

extern "C" int gnntf_spmm_plan_count(const int32_t* row_ptr, int64_t n_rows, int32_t long_threshold,
                                     int32_t chunk, int32_t* counts, void* stream) {
    if (counts == nullptr || (n_rows > 0 && row_ptr == nullptr)) return GNNTF_E_NULL;
    if (n_rows < 0 || n_rows > 0x7ffffffeLL || long_threshold <= 0 || chunk <= 0) return GNNTF_E_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    GNNTF_CUDA_TRY(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), st));
    if (n_rows == 0) return GNNTF_OK;
    const int grid = (int)std::min<int64_t>(ceil_div(n_rows, 256), (int64_t)kNumSMs * 8);
    plan_count_kernel<<<grid, 256, 0, st>>>(row_ptr, (int)n_rows, long_threshold, chunk, counts);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

extern "C" int gnntf_spmm_plan_fill(const int32_t* row_ptr, int64_t n_rows, int32_t long_threshold,
                                    int32_t chunk, int32_t* counters_ws, int32_t* long_row,
                                    int32_t* long_first_chunk, int32_t* long_n_chunks,
                                    int32_t* chunk_row, int32_t* chunk_begin, void* stream) {
    if (counters_ws == nullptr || (n_rows > 0 && row_ptr == nullptr)) return GNNTF_E_NULL;
    if (n_rows < 0 || n_rows > 0x7ffffffeLL || long_threshold <= 0 || chunk <= 0) return GNNTF_E_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    GNNTF_CUDA_TRY(cudaMemsetAsync(counters_ws, 0, 2 * sizeof(int32_t), st));
    if (n_rows == 0) return GNNTF_OK;
    const int grid = (int)std::min<int64_t>(ceil_div(n_rows, 256), (int64_t)kNumSMs * 8);
    plan_fill_kernel<<<grid, 256, 0, st>>>(row_ptr, (int)n_rows, long_threshold, chunk, counters_ws,
                                           long_row, long_first_chunk, long_n_chunks, chunk_row,
                                           chunk_begin);
    GNNTF_LAUNCH_CHECK();
    return GNNTF_OK;
}

extern "C" int gnntf_spmm_f32(const gnntf_csr_t* A, const float* B, int64_t ldb, float* C,
                              int64_t ldc, int64_t F, void* stream) {
    if (C == nullptr && A != nullptr && A->n_rows > 0 && F > 0) return GNNTF_E_NULL;
    Epilogue e{};
    e.s = 1.0f;
    e.C = C;
    e.ldc = ldc;
    e.F = (int)F;
    e.act = GNNTF_ACT_IDENTITY;
    if (F < 0 || F > 0x7fffffff) return GNNTF_E_SIZE;
    return spmm_dispatch(A, B, ldb, e, (cudaStream_t)stream);
}

extern "C" int gnntf_spmm_acc_f32(const gnntf_csr_t* A, const float* B, int64_t ldb, float* C,
                                  int64_t ldc, int64_t F, double scale, void* stream) {
    if (C == nullptr && A != nullptr && A->n_rows > 0 && F > 0) return GNNTF_E_NULL;
    if (F < 0 || F > 0x7fffffff) return GNNTF_E_SIZE;
    Epilogue e{};
    e.s = (float)scale;
    e.act = GNNTF_ACT_IDENTITY;
    e.F = (int)F;
    e.C = nullptr;   // the product only feeds the accumulator
    e.ACC = C;
    e.ldacc = ldc;
    e.u = 0.0f;
    e.w = 1.0f;
    e.acc_init = 0;
    return spmm_dispatch(A, B, ldb, e, (cudaStream_t)stream);
}
