// tc_tabgemm -- "apply a small dense table along one axis of a channels-innermost tensor" on tcgen05:
//
//     Out[g][m][c] = sum_k tab[m][k] * f(In[g][k][c])          g < G groups, m < Mt, k < Kt, c < C
//
// This one kernel is the truncated DFT along y (ydft: g = grid row, k = y, m = (re|im, ky)), the
// truncated C2C DFT along x (xdft: g = (sample, ky), k = (x, re|im), m = (re|im, kx)) and its inverse
// (xsyn), forward and backward -- they only differ in the strides of g, k and m (two-level strides
// cover the (re|im, kx) / (sample, ky) index pairs).
//
// Formulation: channels are the MMA M dimension (128 TMEM lanes; C < 128 packs 128/C groups side
// by side, C > 128 splits into 128-channel blocks), the table rows are N, k is K:
//     D[lane = (g_local, c)][m] = sum_k A[lane][k] * tab[m][k]
//   A : activations, converted in REGISTERS (lane = channel; optional GELU, hi/lo TF32 split) and written to TMEM
//       with tcgen05.st.  They reach the registers through a TMA-fed shared-memory ring of [32 k][128 lanes] chunks
//       (a_tma; one 5-D tensor map covers the two-level strides of g and k): TMEM has room for only two to four A
//       buffers next to the accumulators, and with global loads straight into registers that hand-over depth was also
//       the number of chunks in flight from HBM (xdft / xsyn ran latency bound at 2 TB/s); the ring keeps up to six
//       chunks (96 KB) in flight per SM independently of it.  Shapes TMA cannot address fall back to coalesced
//       128-byte warp loads;
//   B : table slabs [m][32 k] in the canonical K-major SWIZZLE_128B layout, hi and lo parts, either
//       resident for the whole kernel (ydft) or streamed through a ring (xdft / xsyn);
//   D : fp32 accumulator in TMEM; epilogue warps tcgen05.ld it and store 128 contiguous bytes per
//       warp and table row.
// fp32 parity mode is 3xTF32 (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo); FNM_MODE_TF32 issues A_hi*B_hi only.
//
// Warp roles (640 threads, persistent, one CTA per SM; 96 registers at launch, setmaxnreg gives the A loaders
// 128 from what the other warpgroups release):
//   warps 0-3   epilogue                         warp 4       TMEM allocator + MMA issuer
//   warps 5-7   table loaders (group 0)          warps 8-15   two A-loader warpgroups
//   warps 16-19 table loaders (group 1) when the table is streamed, otherwise idle register donors
#include "fnm_tc_ptx.cuh"

#include <cstdio>
#include <cstdlib>

namespace fnm {
namespace tc {

constexpr int kTgNwg = 2;                                 // A-loader warpgroups, two chunks each: needs nbuf >= 4
constexpr int kTgTabWarps = 3;
constexpr int kTgFirstA = 5 + kTgTabWarps;                // first A-loader warp (multiple of 4: TMEM lane quarters)
constexpr int kTgThreads = (kTgFirstA + 4 * kTgNwg + 4) * 32;   // 20 warps: + table group 1 (streamed table) / register donors
constexpr int kTgMaxSlots = 64;
constexpr int kTgMaxBuf = 8;
constexpr uint32_t kTgSmemBudget = 200 * 1024;

struct TgArgs {
    const float* in; const float* tab; float* out;
    long long gin_s1, gin_s0, k_s1, k_s0, gout_s1, gout_s0, m_s1, m_s0;
    int gin_div, kin_div, gout_div, mout_div;
    int C, G, GL, CB, items;
    int Kt, Mt, nchunks, NT, ntiles, dbl, nbuf, reuse, resident, slots, nmain, ncorr;
    int act, single_pass, tab_vec4;
    int pairs;                                             // A loaders take pairs of chunks (needs 4 A buffers) instead of single chunks
    int tab_tma;                                           // streamed table: raw slabs arrive by TMA, the table warps only derive the remainder
    int a_tma, a_slots, a_kb;                              // A chunks through a TMA-fed ring of a_slots x 16 KB; k rows per TMA box (32 or 8)
    int lslots;                                            // tab_tma: `slots` counts RAW slabs (TMA targets), the remainder slabs have their own ring
};

constexpr int kTgLoBar = 32;                               // tab_full / tab_empty index of the first remainder-ring barrier

constexpr int kTgMaxARing = 8;
constexpr uint32_t kTgAChunk = 16384;                      // [32 k][128 lanes] fp32

__device__ __forceinline__ float ld_stream1(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(kTgThreads, 1) tabgemm_tc(const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapA,
                                                            const TgArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t slot_bytes = (uint32_t)a.NT * 256u;
    // streamed table by TMA: [slots raw slabs][lslots remainder slabs], NT x 128 bytes each -- a slab in flight from L2 then
    // takes 16 KB, not a whole (hi | lo) slot, and the ring is deep enough to cover the L2 latency (xdft / xsyn were bound
    // by exactly that: one slab ahead per loader group)
    const uint32_t raw_bytes = (uint32_t)a.NT * 128u;
    uint8_t* lo_ring = smem + (size_t)a.slots * raw_bytes;
    uint8_t* aring = a.tab_tma ? lo_ring + (size_t)a.lslots * raw_bytes : smem + (size_t)a.slots * slot_bytes;   // [a_slots][GL][32 k][min(C, 128)] (a_tma)
    uint64_t* bars = reinterpret_cast<uint64_t*>(aring + (size_t)a.a_slots * kTgAChunk);
    uint64_t* tab_full = bars;                                 // [kTgMaxSlots]
    uint64_t* tab_empty = bars + kTgMaxSlots;                  // [kTgMaxSlots]
    uint64_t* tab_raw = bars + 2 * kTgMaxSlots;                // [kTgMaxSlots] TMA transaction barriers (tab_tma)
    uint64_t* a_full = bars + 3 * kTgMaxSlots;                 // [kTgMaxBuf]
    uint64_t* a_empty = a_full + kTgMaxBuf;                    // [kTgMaxBuf]
    uint64_t* d_full = a_empty + kTgMaxBuf;                    // [2]
    uint64_t* d_empty = d_full + 2;                            // [2]
    uint64_t* ar_full = d_empty + 2;                           // [kTgMaxARing] TMA transaction barriers of the A ring
    uint64_t* ar_empty = ar_full + kTgMaxARing;                // [kTgMaxARing] the 4 warps of the consuming warpgroup
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ar_empty + kTgMaxARing);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.slots; ++s) {
            mbar_init(&tab_full[s], a.resident ? kTgTabWarps : ((s & 1) ? 4 : kTgTabWarps));
            mbar_init(&tab_empty[s], 1);
            mbar_init(&tab_raw[s], 1);
        }
        for (int s = 0; s < a.lslots; ++s) { mbar_init(&tab_full[kTgLoBar + s], (s & 1) ? 4 : kTgTabWarps); mbar_init(&tab_empty[kTgLoBar + s], 1); }
        for (int s = 0; s < a.nbuf; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], 4); }
        for (int s = 0; s < a.a_slots; ++s) { mbar_init(&ar_full[s], 1); mbar_init(&ar_empty[s], 4); }
        fence_barrier_init();
        if (a.tab_tma) tma_prefetch_desc(&mapT);
        if (a.a_tma) tma_prefetch_desc(&mapA);
    }
    if (warp == 4) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t dcol0 = (uint32_t)a.nbuf * 64u;

    const int my_items = (a.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp < 4) {
        // =========================================================================== epilogue
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;\n");
        const int L = warp * 32 + lane;
        uint32_t dcount = 0;
        for (int it = 0; it < my_items; ++it) {
            const int item = (int)blockIdx.x + it * (int)gridDim.x;
            int g, c;
            if (a.C >= 128) { g = item / a.CB; c = (item - g * a.CB) * 128 + L; }
            else { g = item * a.GL + L / a.C; c = L % a.C; }
            const bool valid = g < a.G && c < a.C;
            const int gq = valid ? g / a.gout_div : 0, gr = valid ? g - gq * a.gout_div : 0;
            float* obase = a.out + (size_t)gq * a.gout_s1 + (size_t)gr * a.gout_s0 + c;
            const long long m_wrap = a.m_s1 - (long long)a.mout_div * a.m_s0;
            for (int nt = 0; nt < a.ntiles; ++nt, ++dcount) {
                const int dbuf = a.dbl ? (int)(dcount & 1) : 0;
                const uint32_t dpar = a.dbl ? ((dcount >> 1) & 1) : (dcount & 1);
                mbar_wait(&d_full[dbuf], dpar);
                tc_fence_after();
                const uint32_t dstride = (uint32_t)((a.nmain + a.ncorr) * a.NT);  // nmain main accumulators + corrections
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + dcol0 + (uint32_t)dbuf * dstride;
                const int mlim = min(a.NT, a.Mt - nt * a.NT);
                for (int c0 = 0; c0 < mlim; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c0, r);
                    // + the other main accumulator(s) and the separately accumulated correction terms
                    for (int e = 1; e < a.nmain + (a.single_pass ? 0 : a.ncorr); ++e) {
                        uint32_t rc[16];
                        tmem_ld16(taddr + (uint32_t)(e * a.NT) + c0, rc);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(rc[j]));
                    }
                    tmem_ld_wait();
                    if (valid) {
                        // pointer walk over the (two-level) m strides: one add per store, no multiplies
                        const int m0 = nt * a.NT + c0;
                        const int mq = m0 / a.mout_div;
                        int mr = m0 - mq * a.mout_div;
                        float* op = obase + (size_t)mq * a.m_s1 + (size_t)mr * a.m_s0;
                        const int cnt = mlim - c0;
                        if (a.mout_div >= a.Mt) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                if (j < cnt) *op = __uint_as_float(r[j]);
                                op += a.m_s0;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                if (j < cnt) *op = __uint_as_float(r[j]);
                                op += a.m_s0;
                                if (++mr == a.mout_div) { mr = 0; op += m_wrap; }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[dbuf]);
            }
        }
    } else if (warp == 4) {
        // =========================================================================== MMA issuer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;\n");
        const uint32_t sbase = smem_u32(smem);
        const uint32_t idesc = make_idesc(128, a.NT);
        // Ring positions and phases advance incrementally: the issuer is a single instruction stream, and a `q % slots` /
        // `q / slots` with a run-time divisor is a ~25-instruction dependent sequence -- with five of them per chunk the issuer,
        // not the data, paced xdft / xsyn (ncu stall sampling: the issuing warp never waited).
        uint32_t ab = 0, aph = 0;                 // A buffer of the next chunk and its phase
        uint32_t slot = 0, sph = 0;               // streamed table: raw (or hi | lo) slot and phase
        uint32_t ls = 0, lph = 0;                 // tab_tma: remainder slot and phase
        uint32_t dcount = 0;
        const uint32_t nbuf = (uint32_t)a.nbuf, nslots = (uint32_t)a.slots, nls = (uint32_t)a.lslots;
        for (int it = 0; it < my_items; ++it) {
            const uint32_t ab_item = ab, aph_item = aph;
            for (int nt = 0; nt < a.ntiles; ++nt, ++dcount) {
                const int dbuf = a.dbl ? (int)(dcount & 1) : 0;
                const uint32_t dpar = a.dbl ? ((dcount >> 1) & 1) : (dcount & 1);
                mbar_wait(&d_empty[dbuf], dpar ^ 1);
                // the tensor core accumulates with truncation (measured: ~1.9e-8 relative error per MMA added
                // to an accumulator), so the hi*hi products go to their own accumulator and the two
                // correction passes, 2^-11 smaller, to a second one: the long chain is three times shorter
                // With nmain = 2 the chunks alternate between two main accumulators (chains half as long again);
                // the epilogue adds them with round-to-nearest.
                // Short contractions (<= 4 chunks, ncorr = 0) put the corrections into the main accumulator instead: the
                // chain stays under 50 MMAs and the freed columns double-buffer D, so the epilogue of a tile overlaps the
                // MMAs of the next one.
                const uint32_t dstride = (uint32_t)((a.nmain + a.ncorr) * a.NT);
                const uint32_t d0 = tmem_base + dcol0 + (uint32_t)dbuf * dstride, dc = a.ncorr ? d0 + (uint32_t)(a.nmain * a.NT) : d0;
                uint32_t accum = 0, accum0 = 0, accum1 = 0;
                if (a.reuse) { ab = ab_item; aph = aph_item; }        // every table tile re-reads the item's A chunks
                const bool first_use = !a.reuse || nt == 0, last_use = !a.reuse || nt == a.ntiles - 1;
                int krem = a.Kt;
                for (int ch = 0; ch < a.nchunks; ++ch, krem -= 32) {
                    if (first_use) mbar_wait(&a_full[ab], aph);
                    uint32_t bhi, blo;                               // low descriptor words of the hi / remainder slab
                    uint32_t eslot, elslot = 0;
                    if (a.tab_tma) {
                        // raw slab and remainder slab live in separate rings; the remainder's barrier also covers the raw slab
                        // (the converting warps waited for it)
                        mbar_wait(&tab_full[kTgLoBar + ls], lph);
                        bhi = uniform_u32(kmaj_desc_lo(sbase + slot * raw_bytes));
                        blo = uniform_u32(kmaj_desc_lo(sbase + (nslots + ls) * raw_bytes));
                        eslot = slot; elslot = (uint32_t)kTgLoBar + ls;
                        if (++slot == nslots) { slot = 0; sph ^= 1u; }
                        if (++ls == nls) { ls = 0; lph ^= 1u; }
                    } else {
                        if (a.resident) { eslot = (uint32_t)(nt * a.nchunks + ch); mbar_wait(&tab_full[eslot], 0); }
                        else {
                            eslot = slot;
                            mbar_wait(&tab_full[eslot], sph);
                            if (++slot == nslots) { slot = 0; sph ^= 1u; }
                        }
                        const uint32_t sb = sbase + eslot * slot_bytes;
                        bhi = uniform_u32(kmaj_desc_lo(sb)); blo = bhi + (((uint32_t)a.NT * 128u) >> 4);
                    }
                    tc_fence_after();
                    const uint32_t a_hi = uniform_u32(tmem_base + ab * 64u), a_lo = a_hi + 32u;
                    const int nks = krem >= 32 ? 4 : (krem + 7) >> 3;
                    const bool odd = a.nmain == 2 && (ch & 1);
                    const uint32_t dm = uniform_u32(odd ? d0 + (uint32_t)a.NT : d0);
                    uint32_t acc_m = odd ? accum1 : accum0;
                    const uint32_t acc_c = a.ncorr ? accum : 1u;
                    if (nks == 4) {
                        // the common chunk, fully unrolled: every operand is `uniform base + immediate`
                        if (a.single_pass) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) umma_ts_lo(dm, a_hi + ks * 8, bhi + 2 * ks, idesc, ks == 0 ? acc_m : 1u);
                        } else {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                umma_ts_lo(dm, a_hi + ks * 8, bhi + 2 * ks, idesc, ks == 0 ? acc_m : 1u);
                                umma_ts_lo(dc, a_lo + ks * 8, bhi + 2 * ks, idesc, ks == 0 ? acc_c : 1u);
                                umma_ts_lo(dc, a_hi + ks * 8, blo + 2 * ks, idesc, 1u);
                            }
                        }
                    } else {
                        for (int ks = 0; ks < nks; ++ks) {
                            umma_ts_lo(dm, a_hi + ks * 8, bhi + 2 * ks, idesc, acc_m);
                            acc_m = 1;
                            if (!a.single_pass) {
                                umma_ts_lo(dc, a_lo + ks * 8, bhi + 2 * ks, idesc, (ks == 0) ? acc_c : 1u);
                                umma_ts_lo(dc, a_hi + ks * 8, blo + 2 * ks, idesc, 1);
                            }
                        }
                    }
                    accum = 1;
                    if (odd) accum1 = 1; else accum0 = 1;
                    if (last_use) umma_commit(&a_empty[ab]);
                    if (!a.resident) umma_commit(&tab_empty[eslot]);
                    if (a.tab_tma) umma_commit(&tab_empty[elslot]);
                    if (++ab == nbuf) { ab = 0; aph ^= 1u; }
                }
                umma_commit(&d_full[dbuf]);
                __syncwarp();
            }
        }
    } else if (warp < kTgFirstA || warp >= kTgFirstA + 4 * kTgNwg) {
        // =========================================================================== table loaders
        // resident table (ydft): warps 5-7 fill every slab once.  Streamed table (xdft / xsyn): the table is
        // what bounds the kernel, so the second A-loader warpgroup joins in -- group 0 (warps 5-7) takes the
        // even slabs, group 1 (warps 12-15) the odd ones, two slabs are in flight at any time.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;\n");
        const int grp = warp < kTgFirstA ? 0 : 1;
        const int gthreads = grp == 0 ? kTgTabWarps * 32 : 128;
        const int tt = grp == 0 ? (int)threadIdx.x - 5 * 32 : (int)threadIdx.x - (kTgFirstA + 4 * kTgNwg) * 32;
        auto fill = [&](int nt, int ch, uint8_t* dst) {
            uint8_t* hi = dst;
            uint8_t* lo = dst + (size_t)a.NT * 128;
            const int k0 = ch * 32;
            // batches of 8 independent 16-byte loads per thread (the table is L2 resident: latency, not
            // bandwidth, is what a one-load-at-a-time loop would pay)
            for (int u0 = 0; u0 < a.NT * 8; u0 += gthreads * 8) {
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int u = u0 + tt + j * gthreads;
                    const int r = u >> 3, c4 = u & 7;
                    const int m = nt * a.NT + r, k = k0 + c4 * 4;
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (u < a.NT * 8 && m < a.Mt) {
                        const float* src = a.tab + (size_t)m * a.Kt + k;
                        if (a.tab_vec4 && k + 3 < a.Kt) v[j] = __ldg(reinterpret_cast<const float4*>(src));
                        else {
                            if (k + 0 < a.Kt) v[j].x = __ldg(src + 0);
                            if (k + 1 < a.Kt) v[j].y = __ldg(src + 1);
                            if (k + 2 < a.Kt) v[j].z = __ldg(src + 2);
                            if (k + 3 < a.Kt) v[j].w = __ldg(src + 3);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int u = u0 + tt + j * gthreads;
                    if (u < a.NT * 8) store_split4(hi, lo, u >> 3, u & 7, v[j]);
                }
            }
            fence_proxy_async();
            __syncwarp();
        };
        if (a.resident) {
            for (int nt = 0; nt < (grp == 0 ? a.ntiles : 0); ++nt)
                for (int ch = 0; ch < a.nchunks; ++ch) {
                    const int slot = nt * a.nchunks + ch;
                    fill(nt, ch, smem + (size_t)slot * slot_bytes);
                    if (lane == 0) mbar_arrive(&tab_full[slot]);
                }
        } else if (a.tab_tma) {
            // Streamed table by TMA.  The slab sequence of this CTA is q = 0, 1, ... (item, tile, chunk); group g owns the
            // slabs of its parity and the ring slots of its parity.  One thread of the group issues the raw [NT x 32 k] box
            // of slab q + 2*ahead (once the MMAs of the slot's previous tenant have retired), all its threads then turn the
            // raw slab q -- which IS the hi operand -- into the remainder slab, shared memory to shared memory.
            const uint32_t per_item = (uint32_t)(a.ntiles * a.nchunks);
            const uint32_t total = (uint32_t)my_items * per_item;
            const uint32_t ahead = (uint32_t)a.slots / 2u - 1u;            // slabs of this group in flight beyond the current one
            const bool issuer = (grp == 0 ? warp == 5 : warp == kTgFirstA + 4 * kTgNwg) && lane == 0;
            auto issue = [&](uint32_t q) {
                const uint32_t slot = q % (uint32_t)a.slots;
                const uint32_t r = q % per_item;
                const int nt = (int)(r / (uint32_t)a.nchunks), ch = (int)(r - (uint32_t)nt * a.nchunks);
                mbar_wait(&tab_empty[slot], ((q / (uint32_t)a.slots) & 1) ^ 1);
                mbar_expect_tx(&tab_raw[slot], raw_bytes);
                tma_load_2d(&mapT, &tab_raw[slot], smem + (size_t)slot * raw_bytes, ch * 32, nt * a.NT);
            };
            if (issuer)
                for (uint32_t q = (uint32_t)grp; q < total && q < (uint32_t)grp + 2u * ahead; q += 2) issue(q);   // slots empty: no wait
            for (uint32_t q = (uint32_t)grp; q < total; q += 2) {
                const uint32_t slot = q % (uint32_t)a.slots, ls = q % (uint32_t)a.lslots;
                mbar_wait(&tab_raw[slot], (q / (uint32_t)a.slots) & 1);
                mbar_wait(&tab_empty[kTgLoBar + ls], ((q / (uint32_t)a.lslots) & 1) ^ 1);
                const uint32_t raw0 = smem_u32(smem) + slot * raw_bytes, lo0 = smem_u32(lo_ring) + ls * raw_bytes;
                // two 16-byte pieces per thread and pass: the loads of a pass are in flight together
                for (uint32_t off = (uint32_t)tt * 16u; off < raw_bytes; off += (uint32_t)gthreads * 32u) {
                    float4 v, w;
                    const uint32_t off2 = off + (uint32_t)gthreads * 16u;
                    const bool two = off2 < raw_bytes;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(raw0 + off));
                    if (two) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "r"(raw0 + off2));
                    v.x -= __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); v.y -= __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    v.z -= __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); v.w -= __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo0 + off), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                    if (two) {
                        w.x -= __uint_as_float(__float_as_uint(w.x) & 0xffffe000u); w.y -= __uint_as_float(__float_as_uint(w.y) & 0xffffe000u);
                        w.z -= __uint_as_float(__float_as_uint(w.z) & 0xffffe000u); w.w -= __uint_as_float(__float_as_uint(w.w) & 0xffffe000u);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo0 + off2), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w) : "memory");
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tab_full[kTgLoBar + ls]);
                // refill the ring AFTER handing slab q over: the slot wanted is the one slab q - 2 occupied, whose MMAs
                // have usually retired by now, so the conversion above never waits behind this
                if (issuer && q + 2u * ahead < total) issue(q + 2u * ahead);
            }
        } else {
            uint32_t qs = 0;
            for (int it = 0; it < my_items; ++it)
                for (int nt = 0; nt < a.ntiles; ++nt)
                    for (int ch = 0; ch < a.nchunks; ++ch, ++qs) {
                        if ((int)(qs & 1) != grp) continue;       // slots is even: slab parity == slot parity
                        const uint32_t slot = qs % (uint32_t)a.slots;
                        mbar_wait(&tab_empty[slot], ((qs / (uint32_t)a.slots) & 1) ^ 1);
                        fill(nt, ch, smem + (size_t)slot * slot_bytes);
                        if (lane == 0) mbar_arrive(&tab_full[slot]);
                    }
        }
    } else {
        // =========================================================================== A loaders
        asm volatile("setmaxnreg.inc.sync.aligned.u32 128;\n");
        const int wg = (warp - kTgFirstA) >> 2;
        const int q4 = warp & 3;                                  // TMEM lane quarter this warp may access
        const int L = q4 * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const int per_item = (a.reuse ? 1 : a.ntiles) * a.nchunks;
        const uint32_t total = (uint32_t)my_items * (uint32_t)per_item;
        const long long k_wrap = a.k_s1 - (long long)a.kin_div * a.k_s0;

        // chunk q of this CTA's sequence -> 32 streaming loads (lane = channel, register = k)
        auto load = [&](float (&v)[32], uint32_t q) {
            const int it = (int)(q / (uint32_t)per_item);
            const int ch = (int)(q - (uint32_t)it * per_item) % a.nchunks;
            const int item = (int)blockIdx.x + it * (int)gridDim.x;
            int g, c;
            if (a.C >= 128) { g = item / a.CB; c = (item - g * a.CB) * 128 + L; }
            else { g = item * a.GL + L / a.C; c = L % a.C; }
            const bool valid = g < a.G && c < a.C;
            const int gq = valid ? g / a.gin_div : 0, gr = valid ? g - gq * a.gin_div : 0;
            const float* ibase = a.in + (size_t)gq * a.gin_s1 + (size_t)gr * a.gin_s0 + c;
            const int k0 = ch * 32;
            const int kq = k0 / a.kin_div;
            int kr = k0 - kq * a.kin_div;
            const float* p = ibase + (size_t)kq * a.k_s1 + (size_t)kr * a.k_s0;
            const int klim = valid ? a.Kt - k0 : 0;
            if (a.kin_div >= a.Kt) {                              // uniform stride: one add per load
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = j < klim ? ld_stream1(p) : 0.f;
                    p += a.k_s0;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = j < klim ? ld_stream1(p) : 0.f;
                    p += a.k_s0;
                    if (++kr == a.kin_div) { kr = 0; p += k_wrap; }
                }
            }
        };
        auto process = [&](float (&v)[32], uint32_t q) {
            const uint32_t ab = q % (uint32_t)a.nbuf;
            mbar_wait(&a_empty[ab], ((q / (uint32_t)a.nbuf) & 1) ^ 1);
            tc_fence_after();
            if (a.act) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
            }
            const uint32_t tb = lane_base + ab * 64u;
            split_to_tmem16(tb, tb + 32u, v, !a.single_pass);
            split_to_tmem16(tb + 16u, tb + 48u, v + 16, !a.single_pass);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[ab]);
        };

        // each warpgroup takes PAIRS of consecutive chunks: 64 loads per thread are issued in one batch
        // (2 warpgroups x 128 threads x 256 B = 64 KB in flight per SM), then both chunks are converted.
        // (Prefetching across loop iterations instead makes the consumer wait on the scoreboard that also
        // tracks the younger loads, which serialises everything -- measured 1.6x slower.)
        // With a streamed table TMEM only has room for two A buffers (three accumulators): each warpgroup
        // then takes single chunks, alternating, so one warpgroup's loads fly while the other converts.
        float va[32], vb[32];
        if (a.a_tma) {
            // TMA-fed ring.  This warpgroup's chunks, in order: n = 0, 1, ... -> q(n) (single chunks: every other one; pairs:
            // two consecutive ones out of four).  Thread 0 of the warpgroup issues the boxes of chunk n + ahead once the four
            // warps have read chunk n out of the slot it reuses (q(n + ahead) = q(n) + a_slots: the slots are private to a
            // warpgroup), so up to a_slots chunks are in flight per SM whatever the TMEM hand-over depth is.
            const uint32_t ahead = (uint32_t)a.a_slots / 2u;
            auto qof = [&](uint32_t n) { return a.pairs ? 4u * (n >> 1) + 2u * (uint32_t)wg + (n & 1u) : (uint32_t)wg + 2u * n; };
            const bool issuer = q4 == 0 && lane == 0;
            const int Cb = a.C >= 128 ? 128 : a.C;
            auto issue = [&](uint32_t n) {
                const uint32_t q = qof(n);
                if (q >= total) return;
                const uint32_t slot = q % (uint32_t)a.a_slots;
                mbar_wait(&ar_empty[slot], ((q / (uint32_t)a.a_slots) & 1) ^ 1);
                mbar_expect_tx(&ar_full[slot], kTgAChunk);
                const int it = (int)(q / (uint32_t)per_item);
                const int ch = (int)(q - (uint32_t)it * per_item) % a.nchunks;
                const int item = (int)blockIdx.x + it * (int)gridDim.x;
                uint8_t* dst = aring + (size_t)slot * kTgAChunk;
                for (int gl = 0; gl < a.GL; ++gl) {
                    int g, c0;
                    if (a.C >= 128) { g = item / a.CB; c0 = (item - g * a.CB) * 128; }
                    else { g = item * a.GL + gl; c0 = 0; }
                    // groups / k rows past the end are outside the tensor map and read as zeros
                    const int gq = a.gin_div >= a.G ? 0 : g / a.gin_div, gr = a.gin_div >= a.G ? g : g - gq * a.gin_div;
                    for (int kk = 0; kk < 32; kk += a.a_kb) {
                        const int k = ch * 32 + kk;
                        const int kq = a.kin_div >= a.Kt ? 0 : k / a.kin_div, kr = a.kin_div >= a.Kt ? k : k - kq * a.kin_div;
                        tma_load_5d(&mapA, &ar_full[slot], dst + ((size_t)gl * 32 + kk) * Cb * 4, c0, kr, kq, gr, gq);
                    }
                }
            };
            if (issuer)
                for (uint32_t n = 0; n < ahead; ++n) issue(n);
            const int gl_me = a.C >= 128 ? 0 : L / a.C, c_me = a.C >= 128 ? L : L % a.C;
            for (uint32_t n = 0;; ++n) {
                const uint32_t q = qof(n);
                if (q >= total) break;
                const uint32_t slot = q % (uint32_t)a.a_slots;
                mbar_wait(&ar_full[slot], (q / (uint32_t)a.a_slots) & 1);
                const float* src = reinterpret_cast<const float*>(aring + (size_t)slot * kTgAChunk) + (size_t)gl_me * 32 * Cb + c_me;
#pragma unroll
                for (int j = 0; j < 32; ++j) va[j] = src[j * Cb];
                __syncwarp();
                if (lane == 0) mbar_arrive(&ar_empty[slot]);
                if (issuer) issue(n + ahead);
                __syncwarp();                                       // the issuing lane rejoins before the warp-wide TMEM stores
                process(va, q);
            }
        } else if (a.pairs) {
            for (uint32_t q = 2u * (uint32_t)wg; q < total; q += 2u * kTgNwg) {
                load(va, q);
                const bool two = q + 1 < total;
                if (two) load(vb, q + 1);
                process(va, q);
                if (two) process(vb, q + 1);
            }
        } else {
            for (uint32_t q = (uint32_t)wg; q < total; q += kTgNwg) {
                load(va, q);
                process(va, q);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

static bool tg_force_singles() {
    const char* e = getenv("FNM_TG_SINGLES");
    return e && e[0] == '1';
}

// tile / buffer configuration; returns false when the shape is outside what the kernel serves
static bool tg_configure(TgArgs& a) {
    if (a.Kt < 1 || a.Mt < 1 || a.G < 1) return false;
    if (!(a.C == 16 || a.C == 32 || a.C == 64 || (a.C >= 128 && a.C % 128 == 0))) return false;
    a.GL = a.C >= 128 ? 1 : 128 / a.C;
    a.CB = a.C >= 128 ? a.C / 128 : 1;
    a.items = a.C >= 128 ? a.G * a.CB : (a.G + a.GL - 1) / a.GL;
    a.nchunks = (a.Kt + 31) / 32;
    // pass 0 insists that the A chunks of an item are converted once (reuse); pass 1 accepts re-loading
    // them for every table tile
    for (int pass = 0; pass < 2; ++pass)
        for (int ntiles = 1; ntiles <= 16; ++ntiles) {
            const int NT = ((a.Mt + ntiles - 1) / ntiles + 15) / 16 * 16;
            if (NT > 128) continue;
            for (int dbl = 1; dbl >= 0; --dbl) {
                const bool will_stream = (uint64_t)a.nchunks * ntiles * NT * 256u > kTgSmemBudget || a.nchunks * ntiles > kTgMaxSlots;
                const bool singles = will_stream || tg_force_singles();
                const int nwg_a = singles ? 1 : kTgNwg;            // buffers needed: 2 per warpgroup (pairs), or 1 each (single chunks)
                // long contractions split the main chain over two accumulators when TMEM allows (truncation bias)
                int nmain = (a.nchunks >= 8 && 3 * NT * (dbl ? 2 : 1) + 64 * 2 * nwg_a <= 512) ? 2 : 1;
                if (dbl && nmain == 1 && a.nchunks >= 8 && 3 * NT + 64 * 2 * nwg_a <= 512) continue;   // accuracy first: single-buffered D with two chains
                int ncorr = 1;
                int dcols = (nmain + ncorr) * NT * (dbl ? 2 : 1);    // main accumulator(s) + correction accumulator
                if (dbl && a.nchunks <= 4 && (512 - dcols) / 64 < 2 * nwg_a) {
                    ncorr = 0;                                     // short chain: corrections share the main accumulator
                    dcols = nmain * NT * 2;
                }
                int nbuf = (512 - dcols) / 64;
                if (nbuf > kTgMaxBuf) nbuf = kTgMaxBuf;
                if (nbuf < 2 * nwg_a) continue;                   // mbarrier parity safety of the pair schedule
                const bool reuse = ntiles == 1 || a.nchunks <= nbuf;
                if (!reuse && pass == 0) continue;
                const uint32_t slot_bytes = (uint32_t)NT * 256u;
                const int total = a.nchunks * ntiles;
                const bool resident = (uint64_t)total * slot_bytes <= kTgSmemBudget && total <= kTgMaxSlots;
                int slots = total;
                if (!resident) {
                    slots = (int)(kTgSmemBudget / slot_bytes);
                    if (slots > kTgMaxSlots) slots = kTgMaxSlots;
                    slots &= ~1;                                  // even: the two loader groups alternate slabs
                    if (slots < 2) continue;
                }
                a.NT = NT; a.ntiles = ntiles; a.dbl = dbl; a.nbuf = nbuf; a.reuse = reuse ? 1 : 0; a.nmain = nmain; a.ncorr = ncorr;
                a.resident = resident ? 1 : 0; a.slots = slots; a.pairs = (resident && !singles) ? 1 : 0;
                return true;
            }
        }
    return false;
}

static bool tab_tma_enabled() {
    const char* e = getenv("FNM_TAB_NO_TMA");
    return !(e && e[0] == '1');
}

static int tg_launch(TgArgs& a, const char* what, cudaStream_t st) {
    if (!tg_configure(a)) return fail(FNM_ENOTSUP, "%s: shape not served by the tcgen05 table GEMM", what);
    a.tab_vec4 = (a.Kt % 4 == 0 && (reinterpret_cast<uintptr_t>(a.tab) & 15) == 0) ? 1 : 0;
    static std::atomic<uint64_t> attr_set{0};
    int attr_set_dev = 0;
    if (once_needed(attr_set, attr_set_dev)) {
        const cudaError_t e = cudaFuncSetAttribute(tabgemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail(FNM_ECUDA, "%s: smem attribute: %s", what, cudaGetErrorString(e));
        once_done(attr_set, attr_set_dev);
    }
    const size_t bar_bytes = 1024 + (3 * kTgMaxSlots + 2 * kTgMaxBuf + 4 + 2 * kTgMaxARing) * 8 + 16, total_smem = 227u * 1024u;
    // A chunks through a TMA-fed ring when one 5-D tensor map can address them: dims (c, k_lo, k_hi, g_lo, g_hi), chunks of
    // 32 k that never straddle a k_hi step (boxes of 32 or 8 rows), 16-byte aligned base and strides
    CUtensorMap mapA{}, mapT{};
    a.a_tma = 0; a.a_kb = 32; a.a_slots = 0; a.tab_tma = 0; a.lslots = 0;
    const bool k_uniform = a.kin_div >= a.Kt, g_uniform = a.gin_div >= a.G;
    const int kb = (k_uniform || a.kin_div % 32 == 0) ? 32 : (a.kin_div % 8 == 0 ? 8 : 0);
    bool a_ok;
    {
        const char* e = getenv("FNM_TAB_NO_ATMA");
        a_ok = !(e && e[0] == '1') && kb && (reinterpret_cast<uintptr_t>(a.in) & 15) == 0 &&
               (k_uniform || a.Kt % a.kin_div == 0) && (g_uniform || a.G % a.gin_div == 0) &&
               a.k_s0 % 4 == 0 && a.gin_s0 % 4 == 0 && (k_uniform || a.k_s1 % 4 == 0) && (g_uniform || a.gin_s1 % 4 == 0);
    }
    // streamed table: TMA the raw slabs when the rows are 16-byte aligned.  Ring of raw slabs (as many as fit next to four
    // remainder slabs and a four-chunk A ring) + ring of remainder slabs
    if (!a.resident && a.Kt % 4 == 0 && (reinterpret_cast<uintptr_t>(a.tab) & 15) == 0 && tab_tma_enabled()) {
        const size_t raw = (size_t)a.NT * 128, want_a = a_ok ? 4 * (size_t)kTgAChunk : 0;
        int rs = (int)((total_smem - bar_bytes - want_a - 4 * raw) / raw);
        rs = (rs > 14 ? 14 : rs) & ~1;
        const cuuint64_t dims[2] = {(cuuint64_t)a.Kt, (cuuint64_t)a.Mt};
        const cuuint64_t str[1] = {(cuuint64_t)a.Kt * 4};
        const cuuint32_t box[2] = {32, (cuuint32_t)a.NT};
        if (rs >= 4 && tc_make_map(&mapT, a.tab, 2, dims, str, box, true)) { a.tab_tma = 1; a.slots = rs; a.lslots = 4; }
    }
    const size_t tab_bytes = a.tab_tma ? (size_t)(a.slots + a.lslots) * a.NT * 128 : (size_t)a.slots * a.NT * 256;
    if (a_ok) {
        const cuuint64_t any = (cuuint64_t)a.C * 4;               // stride of an extent-1 dimension (never used)
        const cuuint64_t dims[5] = {(cuuint64_t)a.C, (cuuint64_t)(k_uniform ? a.Kt : a.kin_div), (cuuint64_t)(k_uniform ? 1 : a.Kt / a.kin_div),
                                    (cuuint64_t)(g_uniform ? a.G : a.gin_div), (cuuint64_t)(g_uniform ? 1 : a.G / a.gin_div)};
        const cuuint64_t str[4] = {(cuuint64_t)a.k_s0 * 4, k_uniform ? any : (cuuint64_t)a.k_s1 * 4, (cuuint64_t)a.gin_s0 * 4,
                                   g_uniform ? any : (cuuint64_t)a.gin_s1 * 4};
        const cuuint32_t box[5] = {(cuuint32_t)(a.C >= 128 ? 128 : a.C), (cuuint32_t)kb, 1, 1, 1};
        // the ring takes what the table leaves of the shared memory: 4 to 8 chunks (pairs: whole groups of 4)
        int as = tab_bytes + bar_bytes < total_smem ? (int)((total_smem - tab_bytes - bar_bytes) / kTgAChunk) : 0;
        as = as > 6 && !a.pairs ? 6 : as;
        as = a.pairs ? (as >= 8 ? 8 : (as >= 4 ? 4 : 0)) : (as & ~1);
        if (as >= 4 && tc_make_map_swz(&mapA, a.in, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) { a.a_tma = 1; a.a_kb = kb; a.a_slots = as; }
    }
    size_t smem = tab_bytes + (size_t)a.a_slots * kTgAChunk + bar_bytes;
    if (smem < 120 * 1024) smem = 120 * 1024;                    // one CTA per SM (each owns all 512 TMEM columns)
    {
        const char* e = getenv("FNM_TG_DEBUG");
        if (e && e[0] == '1')
            fprintf(stderr, "[tabgemm %s] C=%d G=%d Kt=%d Mt=%d NT=%d ntiles=%d nchunks=%d nbuf=%d nmain=%d ncorr=%d dbl=%d resident=%d slots=%d lslots=%d "
                    "pairs=%d tab_tma=%d a_tma=%d a_slots=%d a_kb=%d smem=%zu\n", what, a.C, a.G, a.Kt, a.Mt, a.NT, a.ntiles, a.nchunks, a.nbuf,
                    a.nmain, a.ncorr, a.dbl, a.resident, a.slots, a.lslots, a.pairs, a.tab_tma, a.a_tma, a.a_slots, a.a_kb, smem);
    }
    const int grid = a.items < sm_count() ? a.items : sm_count();
    {
        LaunchScope ls(what, st);
        tabgemm_tc<<<grid, kTgThreads, smem, st>>>(mapT, mapA, a);
    }
    return check_launch(what);
}

static bool tg_probe(int C, int G, int Kt, int Mt) {
    if (!tc_enabled()) return false;
    TgArgs a{};
    a.C = C; a.G = G; a.Kt = Kt; a.Mt = Mt;
    return tg_configure(a);
}

}  // namespace tc

using namespace tc;

bool tc_ydft_supported(int64_t rows, int P2, int LY, int C) {
    return rows > 0 && rows < (1LL << 30) && tg_probe(C, (int)rows, P2, LY);
}

int tc_ydft(const float* x, const float* tab, float* T, int64_t rows, int P2, int LY, int C, int act, int mode,
            cudaStream_t st) {
    TgArgs a{};
    a.in = x; a.tab = tab; a.out = T;
    a.C = C; a.G = (int)rows; a.Kt = P2; a.Mt = LY;
    a.gin_div = 1 << 30; a.gin_s1 = 0; a.gin_s0 = (long long)P2 * C;
    a.kin_div = 1 << 30; a.k_s1 = 0; a.k_s0 = C;
    a.gout_div = 1 << 30; a.gout_s1 = 0; a.gout_s0 = (long long)LY * C;
    a.mout_div = 1 << 30; a.m_s1 = 0; a.m_s0 = C;
    a.act = act; a.single_pass = mode == FNM_MODE_TF32;
    return tg_launch(a, "ydft", st);
}

// Xh[r][l][ro][b][c] = sum_{x,ri} ex[(ro,r)][(x,ri)] T[b][x][ri][l][c]:  g = (b,l), k = (x,ri), m = (ro,r)
bool tc_xdft_supported(int B, int P1, int R, int NY, int C) { return tg_probe(C, B * NY, 2 * P1, 2 * R); }

int tc_xdft(const float* T, const float* ex, float* Xh, int B, int P1, int R, int NY, int C, int mode, cudaStream_t st) {
    TgArgs a{};
    a.in = T; a.tab = ex; a.out = Xh;
    a.C = C; a.G = B * NY; a.Kt = 2 * P1; a.Mt = 2 * R;
    a.gin_div = NY; a.gin_s1 = (long long)P1 * 2 * NY * C; a.gin_s0 = C;          // g = b*NY + l
    a.kin_div = 1 << 30; a.k_s1 = 0; a.k_s0 = (long long)NY * C;
    a.gout_div = NY; a.gout_s1 = C; a.gout_s0 = (long long)2 * B * C;
    a.mout_div = R; a.m_s1 = (long long)B * C; a.m_s0 = (long long)NY * 2 * B * C;   // m = ro*R + r
    a.act = 0; a.single_pass = mode == FNM_MODE_TF32;
    return tg_launch(a, "xdft", st);
}

// Z[b][x][ro][l][c] = sum_{ri,r} fx[(x,ro)][(ri,r)] Oh[r][l][ri][b][c]:  g = (b,l), k = (ri,r), m = (x,ro)
bool tc_xsyn_supported(int B, int S1, int R, int NY, int C) { return tg_probe(C, B * NY, 2 * R, 2 * S1); }

int tc_xsyn(const float* Oh, const float* fx, float* Z, int B, int S1, int R, int NY, int C, int mode, cudaStream_t st) {
    TgArgs a{};
    a.in = Oh; a.tab = fx; a.out = Z;
    a.C = C; a.G = B * NY; a.Kt = 2 * R; a.Mt = 2 * S1;
    a.gin_div = NY; a.gin_s1 = C; a.gin_s0 = (long long)2 * B * C;                 // g = b*NY + l
    a.kin_div = R; a.k_s1 = (long long)B * C; a.k_s0 = (long long)NY * 2 * B * C;   // k = ri*R + r
    a.gout_div = NY; a.gout_s1 = (long long)S1 * 2 * NY * C; a.gout_s0 = C;
    a.mout_div = 1 << 30; a.m_s1 = 0; a.m_s0 = (long long)NY * C;
    a.act = 0; a.single_pass = mode == FNM_MODE_TF32;
    return tg_launch(a, "xsyn", st);
}

// out[g][m][j][c] = sum_k tab[m][k] in[g][k][j][c]:  item = (g, j), the table along one axis of a
// channels-last tensor whose trailing extent is inner*C
bool tc_resample_supported(int64_t G, int K, int M, int inner, int C) {
    return G > 0 && inner > 0 && G * inner < (1LL << 30) && tg_probe(C, (int)(G * inner), K, M);
}

int tc_resample(const float* in, const float* tab, float* out, int64_t G, int K, int M, int inner, int C, int mode,
                cudaStream_t st) {
    TgArgs a{};
    a.in = in; a.tab = tab; a.out = out;
    a.C = C; a.G = (int)(G * inner); a.Kt = K; a.Mt = M;
    a.gin_div = inner; a.gin_s1 = (long long)K * inner * C; a.gin_s0 = C;           // item = g*inner + j
    a.kin_div = 1 << 30; a.k_s1 = 0; a.k_s0 = (long long)inner * C;
    a.gout_div = inner; a.gout_s1 = (long long)M * inner * C; a.gout_s0 = C;
    a.mout_div = 1 << 30; a.m_s1 = 0; a.m_s0 = (long long)inner * C;
    a.act = 0; a.single_pass = mode == FNM_MODE_TF32;
    return tg_launch(a, "resample", st);
}

}  // namespace fnm
