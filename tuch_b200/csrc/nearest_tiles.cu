// Geodesically-masked nearest vertex with bounding-sphere pruning.
//
// nearest_kernel (contact_kernels.cu) evaluates all V^2 masked distances of losses.py:92-93.  Here the
// vertices are grouped into the vertex tiles of clusters.cu (32 neighbouring vertices each, ascending
// ids inside a tile), so the 32 candidate rows of a tile -- exactly one word of the bit-packed mask --
// have a small bounding sphere that is recomputed per body.  One warp owns the 32 query columns of one
// tile (lane = query):
//   pass 1: the nearest tile GROUP with an unmasked row (a static per-group summary of the mask), then its tile
//           with the smallest |q - c_t| + R_t; the (few distinct) such tiles of the warp are evaluated first,
//           which gives every query a tight running minimum `best`
//   pass 2: a tile survives for a query iff its mask word is non-zero and (|q - c_t| - R_t)^2 <= best
//           (with fp32 slack that covers the rounding of the expansion-form distance); the warp evaluates
//           the 32 candidates of every tile that survives for ANY of its queries.
// The result is bit-identical to nearest_kernel: same fp32 expansion (|v_r|^2 + |v_c|^2) - 2 v_r.v_c with
// the same FMA order, minimum over a superset of the candidates that can attain it, lowest ORIGINAL row
// index on ties, (0, +inf) for a fully masked column.
#include "api_internal.h"
#include "clusters.h"

#include <cstdlib>

namespace tuch {

constexpr int NT_WARPS = 4;
// first search radius of an interior query in the mixed kernel (metres); see nearest_tiles_kernel<MIXED>
constexpr float NN_INTERIOR_FIRST = 0.06f;

// maskT [W][Vq] (original ids) -> maskP [T][T * 32] over tile slots: bit k of maskP[t][s] is
// geomask[vtile[32 t + k]][vtile[s]]; padding slots are 0
__global__ void permute_mask_kernel(const uint32_t* __restrict__ maskT, int Vq, const int* __restrict__ vtile, int T,
                                    uint32_t* __restrict__ maskP) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (s >= T * 32) return;
    uint32_t bits = 0;
    const int oc = vtile[s];
    if (oc >= 0) {
        for (int k = 0; k < 32; ++k) {
            const int r = vtile[t * 32 + k];
            if (r < 0) break;                                     // padding only at the end of a tile
            bits |= ((maskT[(size_t)(r >> 5) * Vq + oc] >> (r & 31)) & 1u) << k;
        }
    }
    maskP[(size_t)t * T * 32 + s] = bits;
}

// maskG [GW][T * 32], GW = ceil(NG / 32): bit (g & 31) of maskG[g >> 5][s] = group g holds at least one row that
// is unmasked for column s (static per topology, like maskP)
__global__ void group_mask_kernel(const uint32_t* __restrict__ maskP, const int* __restrict__ vgroup_off, int T,
                                  int NG, uint32_t* __restrict__ maskG) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int gw = blockIdx.y;
    if (s >= T * 32) return;
    uint32_t bits = 0;
    for (int k = 0; k < 32 && gw * 32 + k < NG; ++k) {
        const int g = gw * 32 + k;
        uint32_t any = 0;
        for (int t = vgroup_off[g]; t < vgroup_off[g + 1]; ++t) any |= maskP[(size_t)t * T * 32 + s];
        bits |= (any != 0u ? 1u : 0u) << k;
    }
    maskG[(size_t)gw * T * 32 + s] = bits;
}

// one warp per vertex tile:
//   vert4p[b][s] = (v, |v|^2) of vertex vtile[s]  (|v|^2 accumulated exactly like pack_mesh_kernel)
//   tinfo[b][t]  = (centre, radius), (max |v|^2, 0, 0, 0); group spheres follow at [T, T + NG)
__global__ void __launch_bounds__(128)
pack_tiles_kernel(const float* __restrict__ verts, int V, const int* __restrict__ vtile, int T, int NG,
                  float4* __restrict__ vert4p, float4* __restrict__ tinfo) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    const int s = t * 32 + lane;
    const int id = vtile[s];
    const bool ok = id >= 0;
    float x = 0.f, y = 0.f, z = 0.f, w = 0.f;
    if (ok) {
        const float* p = verts + ((size_t)b * V + id) * 3;
        x = p[0]; y = p[1]; z = p[2];
        w = fmaf(z, z, fmaf(y, y, x * x));
    }
    vert4p[(size_t)b * T * 32 + s] = make_float4(x, y, z, w);
    float n = ok ? 1.f : 0.f, sx = x, sy = y, sz = z, wm = w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
        wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    }
    const float inv = 1.f / fmaxf(n, 1.f);
    const float cx = sx * inv, cy = sy * inv, cz = sz * inv;
    float r2 = ok ? (x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
    if (lane == 0) {
        float4* o = tinfo + ((size_t)b * (T + NG) + t) * 2;
        o[0] = make_float4(cx, cy, cz, sqrtf(r2) * 1.0001f + 1e-7f);
        o[1] = make_float4(wm, 0.f, 0.f, 0.f);
    }
}

// tile_any [T][TW], TW = ceil(T / 32): bit (t & 31) of tile_any[qt][t >> 5] = some query slot of tile qt has an
// unmasked row in tile t.  Static per topology (6.5 KB for SMPL).  The tiles next to a query tile on the surface are
// the ones its queries' spheres reach first -- and they are geodesically near, i.e. masked out for the whole tile:
// this bit drops them before any per-query work or mask word is touched.  One warp per (qt, word), lane = tile t.
__global__ void tile_any_kernel(const uint32_t* __restrict__ maskP, int T, uint32_t* __restrict__ tile_any) {
    const int qt = blockIdx.y, w = blockIdx.x, lane = threadIdx.x;
    const int t = w * 32 + lane;
    uint32_t any = 0u;
    if (t < T)
        for (int k = 0; k < 32; ++k) any |= maskP[(size_t)t * T * 32 + qt * 32 + k];
    const unsigned bits = __ballot_sync(0xffffffffu, any != 0u);
    if (lane == 0) tile_any[(size_t)qt * gridDim.x + w] = bits;
}

// one warp per tile group: bounding sphere of its tile spheres and the largest |v|^2
__global__ void __launch_bounds__(128)
pack_groups_kernel(const int* __restrict__ vgroup_off, int T, int NG, float4* __restrict__ tinfo) {
    const int b = blockIdx.y;
    const int g = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= NG) return;
    float4* ib = tinfo + (size_t)b * (T + NG) * 2;
    const int t0 = vgroup_off[g], t1 = vgroup_off[g + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f, wm = 0.f;
    for (int t = t0 + lane; t < t1; t += 32) {
        const float4 s = ib[2 * t];
        sx += s.x; sy += s.y; sz += s.z;
        wm = fmaxf(wm, ib[2 * t + 1].x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
        wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    }
    const float inv = 1.f / (float)(t1 - t0);
    const float cx = sx * inv, cy = sy * inv, cz = sz * inv;
    float r = 0.f;
    for (int t = t0 + lane; t < t1; t += 32) {
        const float4 s = ib[2 * t];
        r = fmaxf(r, sqrtf((s.x - cx) * (s.x - cx) + (s.y - cy) * (s.y - cy) + (s.z - cz) * (s.z - cz)) + s.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    if (lane == 0) {
        ib[2 * (T + g)] = make_float4(cx, cy, cz, r * 1.0001f + 1e-7f);
        ib[2 * (T + g) + 1] = make_float4(wm, 0.f, 0.f, 0.f);
    }
}

// fp32 expansion-form squared distance of contact.py:42, in the FMA order of nearest_kernel
__device__ __forceinline__ float row_value(const float4& v, const float4& q) {
    const float zz = fmaf(v.z, q.z, fmaf(v.y, q.y, v.x * q.x));
    return fmaf(-2.f, zz, v.w + q.w);
}

// smallest unmasked value among the 32 candidate rows of a tile for this lane's query.  Only the VALUE is
// tracked in the hot loop (one FMNMX per row instead of compare + two selects); the row that attains the final
// minimum is looked up once per query at the end (tile_first_row).  FULL: every lane's mask word is all ones.
template <bool FULL>
__device__ __forceinline__ float tile_min(const float4* __restrict__ tv, uint32_t m, const float4& q) {
    float lb = INFINITY;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        float p = row_value(__ldg(tv + k), q);
        if (!FULL) p = ((m >> k) & 1u) ? p : INFINITY;
        lb = fminf(lb, p);
    }
    return lb;
}

// original id of the first row of the tile (rows are in ascending original order) whose value equals target
__device__ __forceinline__ int tile_first_row(const float4* __restrict__ tv, const int* __restrict__ tid, uint32_t m,
                                              const float4& q, float target) {
    int lk = 0;
#pragma unroll 8
    for (int k = 31; k >= 0; --k) {
        const float p = row_value(__ldg(tv + k), q);
        if (((m >> k) & 1u) && p == target) lk = k;
    }
    return __ldg(tid + lk);
}

// running minimum of one query: value, the tile that attains it, and the attaining row's original id once it
// had to be resolved (-1 = not yet).  Exact ties between two tiles go to the lowest original id, as in
// nearest_kernel; they are rare, so both tiles are re-scanned on the spot.
struct NearestState {
    float best = INFINITY;
    int btile = -1;
    int bi = -1;
};

__device__ __forceinline__ void nearest_visit_tile(const float4* __restrict__ vb, const int* __restrict__ vtile,
                                                   const uint32_t* __restrict__ mcol, size_t mstride, int t, uint32_t m,
                                                   bool pad, const float4& q, NearestState& st) {
    const bool full = __all_sync(0xffffffffu, m == 0xffffffffu || pad);
    const float lb = full ? tile_min<true>(vb + t * 32, m, q) : tile_min<false>(vb + t * 32, m, q);
    const bool tie = lb == st.best && lb < INFINITY && t != st.btile && st.btile >= 0;
    if (lb < st.best) { st.best = lb; st.btile = t; st.bi = -1; }
    if (__any_sync(0xffffffffu, tie)) {
        if (tie) {
            if (st.bi < 0)
                st.bi = tile_first_row(vb + st.btile * 32, vtile + st.btile * 32, mcol[(size_t)st.btile * mstride], q, st.best);
            const int r = tile_first_row(vb + t * 32, vtile + t * 32, m, q, lb);
            if (r < st.bi) { st.bi = r; st.btile = t; }
        }
    }
}

// grid (groups of NT_WARPS query tiles, bodies)
//
// MIXED: an exterior query only counts candidates within `limit` (metres); an interior one has no limit.
// SMPLify-DC's contact term consumes the nearest allowed vertex of an INTERIOR vertex at any distance, but
// of an EXTERIOR vertex only when it is closer than euclthres (losses.py:96-103, 2 cm).  In this mode the kernel is a
// radius search for every query (stage A below: `limit` for the exterior ones, interior_first for the interior
// ones) and the unlimited search moves out: the interior queries without an allowed vertex inside their first
// radius are appended to todo_list for nearest_single_kernel.  A query with an allowed vertex inside its radius gets
// exactly the unlimited answer (same values, same tie-breaking); an exterior one without gets (-1, +inf) -- or
// (0, +inf), as in the unlimited case, when its mask column is empty.  The radius carries the slack by which an fp32
// expansion-form value can exceed the true squared distance.
// exterior (MIXED only, [B][V]): the inside test's flags; 0 = interior.
template <bool MIXED>
__global__ void __launch_bounds__(NT_WARPS * 32)
nearest_tiles_kernel(const float4* __restrict__ vert4p, const float4* __restrict__ tinfo,
                     const uint32_t* __restrict__ maskP, const uint32_t* __restrict__ maskG,
                     const uint32_t* __restrict__ tile_any, const int* __restrict__ vtile,
                     const int* __restrict__ vgroup_off, int V, int T, int NG, float limit, float interior_first,
                     const uint8_t* __restrict__ exterior, int* __restrict__ todo_list,
                     int* __restrict__ argmin_out, float* __restrict__ min_out) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int qt = blockIdx.x * NT_WARPS + (threadIdx.x >> 5);
    if (qt >= T) return;
    const int slot = qt * 32 + lane;                                   // query column (tile slot)
    const int oc = vtile[slot];                                        // original vertex id, -1 = padding
    const bool limited = MIXED && oc >= 0 && exterior[(size_t)b * V + oc] != 0;
    const float4* vb = vert4p + (size_t)b * T * 32;
    const float4* ib = tinfo + (size_t)b * (T + NG) * 2;
    const float4 q = vb[slot];
    const uint32_t* mcol = maskP + slot;                               // mask words of this column, stride T*32
    const size_t mstride = (size_t)T * 32;
    const bool pad = oc < 0;
    const uint32_t* tany = tile_any + (size_t)qt * ((T + 31) >> 5);    // which tiles hold any allowed row for this tile
    NearestState st;
    bool seeded = oc >= 0;                                             // takes part in the unlimited search below

    if (MIXED) {
        // Stage A: EVERY query searches within a radius first -- `limit` for the exterior ones (all they need), and
        // NN_INTERIOR_FIRST for the interior ones, whose nearest allowed vertex is almost always the surface they
        // have sunk into, a few centimetres away.  All candidates then lie within that radius of this tile's own
        // sphere, so the groups, and the tiles of a reachable group, are tested sphere against sphere with ONE LANE
        // PER GROUP / TILE instead of one loop trip each (29 + ~35 trips of dependent loads); the few survivors get
        // the per-query test of the general loop.  A query with an allowed vertex inside its radius has its exact,
        // final answer after this stage; only interior queries without one go on to the unlimited search.
        const float radius = limited ? limit : fmaxf(limit, interior_first);
        if (oc >= 0) st.best = fmaf(radius * radius, 1.001f, 4e-6f * (3.f * q.w + 2.f * radius * radius));
        float thr = fmaf(st.best, 1.00001f, 4e-6f * q.w);
        const bool any_interior = __any_sync(0xffffffffu, oc >= 0 && !limited);
        const float4 own = __ldg(ib + 2 * qt);
        const float reach_r = own.w + 1.01f * (any_interior ? fmaxf(limit, interior_first) : limit) + 1e-6f;
        for (int g0 = 0; g0 < NG; g0 += 32) {
            bool gcand = false;
            if (g0 + lane < NG) {
                const float4 gs = __ldg(ib + 2 * (T + g0 + lane));
                const float gx = own.x - gs.x, gy = own.y - gs.y, gz = own.z - gs.z;
                gcand = fmaf(sqrtf(fmaf(gz, gz, fmaf(gy, gy, gx * gx))), 0.9999f, -gs.w) <= reach_r;
            }
            for (unsigned gm = __ballot_sync(0xffffffffu, gcand); gm != 0u; gm &= gm - 1u) {
                const int g = g0 + __ffs(gm) - 1;
                const int t0 = __ldg(vgroup_off + g), t1 = __ldg(vgroup_off + g + 1);
                for (int tb = t0; tb < t1; tb += 32) {
                    bool tcand = false;
                    if (tb + lane < t1 && ((__ldg(tany + ((tb + lane) >> 5)) >> ((tb + lane) & 31)) & 1u)) {
                        const float4 s = __ldg(ib + 2 * (tb + lane));
                        const float dx = own.x - s.x, dy = own.y - s.y, dz = own.z - s.z;
                        tcand = fmaf(sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx))), 0.9999f, -s.w) <= reach_r;
                    }
                    for (unsigned tm = __ballot_sync(0xffffffffu, tcand); tm != 0u; tm &= tm - 1u) {
                        const int t = tb + __ffs(tm) - 1;
                        const float4 s = __ldg(ib + 2 * t), s2 = __ldg(ib + 2 * t + 1);
                        const float dx = q.x - s.x, dy = q.y - s.y, dz = q.z - s.z;
                        const float lo = fmaf(sqrt_approx(fmaf(dz, dz, fmaf(dy, dy, dx * dx))), 0.9999f, -s.w);
                        const bool reach = oc >= 0 && (lo <= 0.f || lo * lo <= fmaf(4e-6f, s2.x, thr));
                        if (!__any_sync(0xffffffffu, reach)) continue;
                        const uint32_t m = mcol[(size_t)t * mstride];
                        if (!__any_sync(0xffffffffu, reach && m != 0u)) continue;
                        nearest_visit_tile(vb, vtile, mcol, mstride, t, m, pad, q, st);
                        thr = fmaf(st.best, 1.00001f, 4e-6f * q.w);
                    }
                }
            }
        }
        // Interior queries that found nothing within their first radius (a fifth of them, < 1 % of all queries) are
        // handed to nearest_single_kernel, one WARP per query: as lanes of this kernel each of them kept a whole
        // warp on the serial unlimited search below for ~100 us, which set the duration of the kernel.
        seeded = oc >= 0 && !limited && st.btile < 0;
        const unsigned hand = __ballot_sync(0xffffffffu, seeded);
        if (hand != 0u) {
            int base = 0;
            if (lane == 0) base = atomicAdd(todo_list, __popc(hand));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (seeded) todo_list[1 + base + __popc(hand & ((1u << lane) - 1u))] = b * (T * 32) + slot;
        }
    }

    if (!MIXED) {
    // pass 1 (a heuristic: pass 2 is exhaustive whatever it picks): the nearest group -- by the lower bound of
    // its sphere -- that holds an unmasked row for this query, then the tile of that group whose sphere promises
    // the smallest masked distance |q - c| + R.  Queries of a warp are neighbours: few distinct groups.
    int gstar = -1;
    {
        float glo_best = INFINITY;
        for (int g = 0; g < NG; ++g) {
            const float4 gs = __ldg(ib + 2 * (T + g));
            const float gx = q.x - gs.x, gy = q.y - gs.y, gz = q.z - gs.z;
            const float glo = sqrt_approx(fmaf(gz, gz, fmaf(gy, gy, gx * gx))) - gs.w;
            const bool has = (maskG[(size_t)(g >> 5) * mstride + slot] >> (g & 31)) & 1u;
            if (has && seeded && glo < glo_best) { glo_best = glo; gstar = g; }
        }
    }
    float ub = INFINITY;
    int tstar = -1;
    unsigned gtodo = __ballot_sync(0xffffffffu, gstar >= 0);
    while (gtodo != 0u) {
        const int g = __shfl_sync(0xffffffffu, gstar, __ffs(gtodo) - 1);
        const bool mine = gstar == g;
        const int t0 = __ldg(vgroup_off + g), t1 = __ldg(vgroup_off + g + 1);
        for (int t = t0; t < t1; ++t) {
            const uint32_t m = mcol[(size_t)t * mstride];
            const float4 s = __ldg(ib + 2 * t);
            const float dx = q.x - s.x, dy = q.y - s.y, dz = q.z - s.z;
            const float d = sqrt_approx(fmaf(dz, dz, fmaf(dy, dy, dx * dx))) + s.w;
            if (mine && m != 0u && d < ub) { ub = d; tstar = t; }
        }
        gtodo &= ~__ballot_sync(0xffffffffu, mine);
    }
    // evaluate those tiles first (a warp's queries are neighbours: few distinct ones): from here on
    // `best` is a tight bound for the sphere test
    const bool live = tstar >= 0;                                      // votes in pass 2
    unsigned todo = __ballot_sync(0xffffffffu, tstar >= 0);
    while (todo != 0u) {
        const int ts = __shfl_sync(0xffffffffu, tstar, __ffs(todo) - 1);
        nearest_visit_tile(vb, vtile, mcol, mstride, ts, mcol[(size_t)ts * mstride], pad, q, st);
        todo &= ~__ballot_sync(0xffffffffu, tstar == ts);
    }
    // pass 2: every tile whose sphere can still hold a row at or below the running minimum; groups of
    // tiles are tested first.  The slack covers the rows' fp32 expansion-form values undershooting their
    // true squared distance (<= 3.6e-7 (|v|^2 + |q|^2)).
    // thr: the running minimum with the query's share of the slack folded in, refreshed after every visit
    float thr = fmaf(st.best, 1.00001f, 4e-6f * q.w);
    for (int g = 0; g < NG; ++g) {
        const float4 gs = __ldg(ib + 2 * (T + g)), gs2 = __ldg(ib + 2 * (T + g) + 1);
        const float gx = q.x - gs.x, gy = q.y - gs.y, gz = q.z - gs.z;
        const float glo = fmaf(sqrt_approx(fmaf(gz, gz, fmaf(gy, gy, gx * gx))), 0.9999f, -gs.w);
        // lanes without any unmasked row (padding slots, fully masked columns) never vote
        const bool gneed = live && (glo <= 0.f || glo * glo <= fmaf(4e-6f, gs2.x, thr));
        if (!__any_sync(0xffffffffu, gneed)) continue;
        const int t0 = __ldg(vgroup_off + g), t1 = __ldg(vgroup_off + g + 1);
        const uint32_t* mp = mcol + (size_t)t0 * mstride;
        const float4* sp = ib + 2 * t0;
        for (int t = t0; t < t1; ++t, mp += mstride, sp += 2) {
            if (!((__ldg(tany + (t >> 5)) >> (t & 31)) & 1u)) continue;   // masked out for the whole query tile
            // requested up front: its L2 round trip overlaps the sphere test.  (Tried for the mixed kernel, whose few
            // warps here run at low occupancy: a group's eight mask words requested together -- 70 registers,
            // 447 against 363 us.)
            const uint32_t m = *mp;
            const float4 s = __ldg(sp), s2 = __ldg(sp + 1);
            const float dx = q.x - s.x, dy = q.y - s.y, dz = q.z - s.z;
            const float lo = fmaf(sqrt_approx(fmaf(dz, dz, fmaf(dy, dy, dx * dx))), 0.9999f, -s.w);
            const bool need = m != 0u && (!MIXED || live) && (lo <= 0.f || lo * lo <= fmaf(4e-6f, s2.x, thr));
            if (!__any_sync(0xffffffffu, need)) continue;
            if (__any_sync(0xffffffffu, tstar == t)) continue;          // evaluated for the whole warp above
            nearest_visit_tile(vb, vtile, mcol, mstride, t, m, pad, q, st);
            thr = fmaf(st.best, 1.00001f, 4e-6f * q.w);
        }
    }
    }
    // the attaining rows: every query scans the 32 rows of ITS attaining tile once (per-lane gathers -- a warp's
    // queries attain their minima in ~9 different tiles, and one uniform scan per distinct tile was 17 % of the
    // kernel's instructions)
    if (st.bi < 0 && st.btile >= 0)
        st.bi = tile_first_row(vb + st.btile * 32, vtile + st.btile * 32, mcol[(size_t)st.btile * mstride], q, st.best);
    if (oc >= 0 && !(MIXED && seeded)) {
        const bool none = st.btile < 0;                                // fully masked column, or nothing within the limit
        int none_id = 0;
        if (limited && none) {
            uint32_t any = 0u;
            for (int w = 0; w < (NG + 31) / 32; ++w) any |= maskG[(size_t)w * mstride + slot];
            none_id = any != 0u ? -1 : 0;
        }
        argmin_out[(size_t)b * V + oc] = none ? none_id : st.bi;
        min_out[(size_t)b * V + oc] = none ? INFINITY : st.best;
    }
}

// The unlimited search of ONE query per warp (the interior queries the mixed kernel hands over): lanes are tiles in
// the bounding phase and rows in the evaluation phase, so the dependent chain is ~30 round trips instead of the
// ~150 of a lane of nearest_tiles_kernel.  Same values (row_value), same winner: the smallest value, the lowest
// original row id among equal ones, (0, +inf) for an empty mask column.
// Dynamic shared memory: 2 T words per warp (the tiles' lower bounds and the query's mask words).
__global__ void __launch_bounds__(NT_WARPS * 32)
nearest_single_kernel(const float4* __restrict__ vert4p, const float4* __restrict__ tinfo,
                      const uint32_t* __restrict__ maskP, const int* __restrict__ vtile, int V, int T, int NG,
                      const int* __restrict__ todo_list, int* __restrict__ argmin_out, float* __restrict__ min_out) {
    extern __shared__ float s_lo[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* lo = s_lo + (size_t)warp * 2 * T;
    uint32_t* mw = (uint32_t*)(lo + T);                              // the query's mask word of every tile
    const size_t mstride = (size_t)T * 32;
    const int n = todo_list[0];
    for (int e = blockIdx.x * NT_WARPS + warp; e < n; e += gridDim.x * NT_WARPS) {
        const int id = todo_list[1 + e];
        const int b = id / (T * 32), slot = id - b * (T * 32);
        const float4* vb = vert4p + (size_t)b * T * 32;
        const float4* ib = tinfo + (size_t)b * (T + NG) * 2;
        const float4 q = vb[slot];
        const uint32_t* mcol = maskP + slot;
        // lower bound of every tile that holds an allowed row (scaled like the sphere test of the tile kernel)
        float vmax2 = 0.f;           // largest |v|^2 of any candidate: the slack of the pruning test below
        for (int t = lane; t < T; t += 32) {
            float v = INFINITY;
            const uint32_t m = mcol[(size_t)t * mstride];
            mw[t] = m;
            if (m != 0u) {
                const float4 s = __ldg(ib + 2 * t);
                const float dx = q.x - s.x, dy = q.y - s.y, dz = q.z - s.z;
                v = fmaxf(fmaf(sqrt_approx(fmaf(dz, dz, fmaf(dy, dy, dx * dx))), 0.9999f, -s.w), 0.f);
                vmax2 = fmaxf(vmax2, __ldg(ib + 2 * t + 1).x);
            }
            lo[t] = v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vmax2 = fmaxf(vmax2, __shfl_xor_sync(0xffffffffu, vmax2, o));
        const float slack = 4e-6f * (q.w + vmax2);
        __syncwarp();
        float best = INFINITY;       // of this lane's rows
        int best_id = 0x7fffffff;
        float wbest = INFINITY;      // of the warp: the pruning bound
        while (true) {
            float m = INFINITY;
            int mt = -1;
            for (int t = lane; t < T; t += 32) {
                const float v = lo[t];
                if (v < m) { m = v; mt = t; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float om = __shfl_xor_sync(0xffffffffu, m, o);
                const int ot = __shfl_xor_sync(0xffffffffu, mt, o);
                if (om < m || (om == m && ot >= 0 && (mt < 0 || ot < mt))) { m = om; mt = ot; }
            }
            if (mt < 0 || m == INFINITY) break;
            if (!(m * m <= fmaf(wbest, 1.00001f, slack))) break;       // nor can any other tile: m is the smallest bound
            const uint32_t bits = mw[mt];
            const int rid = vtile[mt * 32 + lane];
            if (((bits >> lane) & 1u) && rid >= 0) {
                const float p = row_value(__ldg(vb + mt * 32 + lane), q);
                if (p < best || (p == best && rid < best_id)) { best = p; best_id = rid; }
            }
            float w = best;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) w = fminf(w, __shfl_xor_sync(0xffffffffu, w, o));
            wbest = w;
            if (lane == 0) lo[mt] = INFINITY;
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_id, o);
            if (ob < best || (ob == best && oi < best_id)) { best = ob; best_id = oi; }
        }
        const int oc = vtile[slot];
        if (lane == 0 && oc >= 0) {
            const bool none = !(best < INFINITY);
            argmin_out[(size_t)b * V + oc] = none ? 0 : best_id;
            min_out[(size_t)b * V + oc] = none ? INFINITY : best;
        }
        __syncwarp();
    }
}

int launch_permute_mask(const uint32_t* maskT, int Vq, const int* vtile, int T, uint32_t* maskP, cudaStream_t st) {
    dim3 grid(cdiv(T * 32, 128), T);
    permute_mask_kernel<<<grid, 128, 0, st>>>(maskT, Vq, vtile, T, maskP);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_tile_any_mask(const uint32_t* maskP, int T, uint32_t* tile_any, cudaStream_t st) {
    dim3 grid(cdiv(T, 32), T);
    tile_any_kernel<<<grid, 32, 0, st>>>(maskP, T, tile_any);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_group_mask(const uint32_t* maskP, const int* vgroup_off, int T, int NG, uint32_t* maskG, cudaStream_t st) {
    dim3 grid(cdiv(T * 32, 128), cdiv(NG, 32));
    group_mask_kernel<<<grid, 128, 0, st>>>(maskP, vgroup_off, T, NG, maskG);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

// tile / group spheres of every body, then the query kernel over the bodies [b0, b0 + nb): the fused iteration
// launches the bodies in two parts, one ahead of the inside test and one behind it (fit_api.cu)
int launch_nearest_tiles_pack(const float* verts, const int* vtile, const int* vgroup_off, int B, int V, int T, int NG,
                              float4* vert4p, float4* tinfo, cudaStream_t st) {
    if (B == 0) return 0;
    KernelTimer timer("nearest_pack_kernels", st);
    dim3 grid(cdiv(T, 4), B);
    pack_tiles_kernel<<<grid, 128, 0, st>>>(verts, V, vtile, T, NG, vert4p, tinfo);
    TUCH_LAUNCH_CHECK(); count_launch();
    dim3 g2(cdiv(NG, 4), B);
    pack_groups_kernel<<<g2, 128, 0, st>>>(vgroup_off, T, NG, tinfo);
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

// limit < 0: every query unlimited; else the queries whose exterior byte is non-zero are limited (see the kernel) and
// todo_list ([1 + nb * V] ints of scratch) takes the interior queries that go on to nearest_single_kernel
int launch_nearest_tiles_query(const uint32_t* maskP, const uint32_t* maskG, const uint32_t* tile_any, const int* vtile,
                               const int* vgroup_off, int b0, int nb, int V, int T, int NG, const float4* vert4p,
                               const float4* tinfo, float limit, const uint8_t* exterior, int* todo_list, int* argmin,
                               float* minval, cudaStream_t st) {
    if (nb == 0) return 0;
    dim3 grid(cdiv(T, NT_WARPS), nb);
    KernelTimer timer("nearest_kernel", st);
    const float4* v4 = vert4p + (size_t)b0 * T * 32;
    const float4* ti = tinfo + (size_t)b0 * (T + NG) * 2;
    static const float first = getenv("TUCH_NN_INTERIOR_FIRST") ? (float)atof(getenv("TUCH_NN_INTERIOR_FIRST")) : NN_INTERIOR_FIRST;  // dev knob
    if (limit >= 0.f && exterior != nullptr) {
        TUCH_REQUIRE(todo_list != nullptr, "nearest vertex: the mixed query needs its hand-over list");
        TUCH_CUDA(cudaMemsetAsync(todo_list, 0, sizeof(int), st));
        nearest_tiles_kernel<true><<<grid, NT_WARPS * 32, 0, st>>>(v4, ti, maskP, maskG, tile_any, vtile, vgroup_off, V, T, NG, limit,
                                                                   first, exterior + (size_t)b0 * V, todo_list,
                                                                   argmin + (size_t)b0 * V, minval + (size_t)b0 * V);
        TUCH_LAUNCH_CHECK(); count_launch();
        const size_t smem = sizeof(float) * 2 * (size_t)T * NT_WARPS;
        TUCH_REQUIRE(smem <= 48 * 1024, "nearest vertex: %d vertex tiles exceed the shared memory of the single-query kernel", T);
        nearest_single_kernel<<<sm_count() * 32, NT_WARPS * 32, smem, st>>>(v4, ti, maskP, vtile, V, T, NG, todo_list,
                                                                            argmin + (size_t)b0 * V, minval + (size_t)b0 * V);
    } else {
        nearest_tiles_kernel<false><<<grid, NT_WARPS * 32, 0, st>>>(v4, ti, maskP, maskG, tile_any, vtile, vgroup_off, V, T, NG, -1.f,
                                                                    0.f, nullptr, nullptr, argmin + (size_t)b0 * V,
                                                                    minval + (size_t)b0 * V);
    }
    TUCH_LAUNCH_CHECK(); count_launch();
    return 0;
}

int launch_nearest_tiles(const float* verts, const uint32_t* maskP, const uint32_t* maskG, const uint32_t* tile_any,
                         const int* vtile,
                         const int* vgroup_off, int B, int V, int T, int NG, float4* vert4p, float4* tinfo,
                         int* argmin, float* minval, cudaStream_t st) {
    if (B == 0) return 0;
    if (int rc = launch_nearest_tiles_pack(verts, vtile, vgroup_off, B, V, T, NG, vert4p, tinfo, st)) return rc;
    return launch_nearest_tiles_query(maskP, maskG, tile_any, vtile, vgroup_off, 0, B, V, T, NG, vert4p, tinfo, -1.f, nullptr, nullptr,
                                      argmin, minval, st);
}

}  // namespace tuch
