// fused.cu -- small-window path: the integral planes never reach HBM.
//
// Same masks as kernel 1 + kernel 2 (integral.cu, threshold.cu) for Sauvola / Niblack / NICK / Feng with
// windows up to 31, at ~1 byte read + 1 byte written per pixel instead of ~35.
//
// Window sums do not need an integral image at all: per column keep the running sum V over the last d rows
// (add the entering row, subtract the leaving one), and take the horizontal window as a difference of the row
// prefix of V.  So:
//   * a page is cut into strips of 128 padded columns; ONE WARP owns a strip and walks down all rows (no block
//     barrier, no cross-warp traffic).  Lane l owns 4 adjacent columns: V_S, V_Q in registers, the last d+1 rows
//     of raw pixels in a 2 KB shared-memory ring (to subtract the leaving row and to fetch the centre pixel);
//   * per row: update V, lane prefix + two 5-step shuffle scans of V, exchange of the prefix through a
//     double-buffered shared row, window sums S_win / Q_win = prefix[x+d] - prefix[x]  (exact integers);
//   * decision in three tiers, each sound on its own (decide.cuh):
//       1. FP32 estimate (+ variance-free bounds), margin mu ~ 2e-3            -> settles ~99.9 % of the pixels
//       2. FP64 estimate from the same exact integers, margin mu2 ~ 1e-5 that bounds the REFERENCE's own FP64
//          rounding (which depends on the absolute int64 integral values, unknown here)
//       3. the few pixels per page that are still undecided are written as the sentinel 128 and finished by
//          fixup_kernel, which rebuilds the four true int64 taps by brute force (row prefixes at the strip start
//          come from the pre-pass strip_rowsum_kernel) and evaluates the reference's FP64 formula literally.
//   Strips overlap by d columns (a strip emits 128 - d outputs): ~12 % redundant work at w = 15.
//
// Measured and dropped (B200, 256 A4 pages, w = 15, masks identical in every case):
//   * 8 columns per lane over strips of 256 columns (half the scans / exchanges / overlap per pixel): 138 registers at
//     3 CTAs per SM 8.25 ms, squeezed to 128 registers at 4 CTAs per SM 7.20 ms -- against 7.31 ms for this kernel, and
//     slower for w = 7, 31 and small pages.  The per-row fixed cost is not what bounds the kernel; the per-pixel decision is.
//   * a per-warp queue of undecided pixels handled after the row loop: 8 % slower (divergence moved, not removed).
#include "common.cuh"
#include "decide.cuh"
#include <algorithm>

namespace {

constexpr int kFW = 128;          // scanned padded columns per warp strip
constexpr int kSentinel = 128;    // mask value of a pixel left to fixup_kernel
constexpr int kPageCap = 128;     // listed pixels per page; a page with more goes to the planes path

struct FusedArgs {
    const uint8_t* src; size_t src_step, src_page_stride;
    uint8_t* dst; size_t dst_step, dst_page_stride;
    const uint32_t* imin;            // per page (Feng)
    uint32_t* blocksum;              // [page][nblk][ns][2] {sum, sqsum} over a block of 32 padded rows x the strip's own ow columns
    uint32_t* scount;                // [page] pixels left to the fixup
    uint32_t* slist;                 // [page][kPageCap] (y << 16) | x
    int rows, cols, pad, d, Hp, Wp, out_rows, out_cols, ow, ns, n_pages;
    uint32_t cap;                    // listed pixels per page the fixup takes (<= kPageCap)
    int nb, band_rows, nblk;         // row bands per page, output rows per band, 32-row blocks per page
    int no_tier2;                    // validation: skip the FP64 estimate, every tier-1 leftover goes to the list
    double kw, p0, p1, p2;
};

// ---- pre-pass (Feng only): per-page minimum
__global__ void __launch_bounds__(256)
page_min_kernel(const uint8_t* __restrict__ src, size_t step, size_t page_stride, int rows, int cols, uint32_t* __restrict__ imin)
{
    const int page = blockIdx.y;
    const uint8_t* img = src + (size_t)page * page_stride;
    uint32_t mn = 255u;
    for (int y = blockIdx.x; y < rows; y += gridDim.x) {
        const uint8_t* row = img + (size_t)y * step;
        for (int x = threadIdx.x; x < cols; x += 256) mn = min(mn, (uint32_t)__ldg(row + x));
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    if ((threadIdx.x & 31) == 0) atomicMin(imin + page, mn);
}

// byte i of w as an integer (one PRMT)
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return __byte_perm(w, 0u, 0x4440u + (uint32_t)i); }

// one step of an inclusive warp scan: SHFL with its in-range predicate + predicated add
__device__ __forceinline__ uint32_t scan_step(uint32_t v, int delta)
{
    uint32_t r;
    asm("{\n\t.reg .u32 t;\n\t.reg .pred p;\n\tshfl.sync.up.b32 t|p, %1, %2, 0x0, 0xffffffff;\n\tmov.u32 %0, %1;\n\t@p add.u32 %0, %0, t;\n\t}"
        : "=r"(r) : "r"(v), "r"(delta));
    return r;
}
__device__ __forceinline__ uint32_t warp_scan_incl(uint32_t v)
{
    v = scan_step(v, 1); v = scan_step(v, 2); v = scan_step(v, 4); v = scan_step(v, 8); v = scan_step(v, 16);
    return v;
}

// ---- the fused kernel: one warp = one strip of one band of one page
template <int METHOD, int RR, bool N32>
__global__ void __launch_bounds__(128, 5)
local_fused_kernel(const FusedArgs A, const FastArgs F)
{
    constexpr int WARPS = 4;
    // per warp: pixel ring [RR][32 words], prefix exchange [2][2 planes][128 words] (+ slack for tap over-reads)
    constexpr int WARP_WORDS = RR * 32 + 2 * 2 * kFW + 16;
    __shared__ __align__(16) uint32_t fsm[WARPS * WARP_WORDS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gw = blockIdx.x * WARPS + wid;
    const int per_page = A.ns * A.nb;
    if (gw >= per_page * A.n_pages) return;             // warps are independent: no block-level barrier below
    const int page = gw / per_page, rem = gw - page * per_page;
    const int band = rem / A.ns, strip = rem - band * A.ns;
    const int x0 = strip * A.ow;                        // first scanned padded column == first output column
    uint32_t* pixr = fsm + wid * WARP_WORDS;
    uint32_t* xch = pixr + RR * 32;
    const uint8_t* src = A.src + (size_t)page * A.src_page_stride;
    const int pad = A.pad, d = A.d;
    // output rows [yo0, yo1) need the padded rows yo0+1 .. yo1-1+d; the first d-1 of them only warm the column sums up
    const int yo0 = band * A.band_rows, yo1 = min(yo0 + A.band_rows, A.out_rows);
    const int Ybeg = band == 0 ? 0 : yo0 + 1, Yend = yo1 - 1 + d;
    const int own0 = band == 0 ? 0 : yo0 + d;           // padded rows this band accounts for in the block sums

    double imin = 0.0;
    float pbh = F.pb_hi, pbl = F.pb_lo;
    if (METHOD == PRL_FENG) {
        imin = (double)A.imin[page];
        const float iminf = (float)imin, c3 = fmaf(F.c2, iminf, -iminf);
        pbh += c3; pbl += c3;
    }
    const float mu = F.mu0;
    const float Hc = pbh + 0.5f + mu, Lc = pbl + 0.5f - mu;

    // lane's padded columns X = x0 + 4*lane + i  <->  source columns xs0 + i
    const int xs0 = x0 + 4 * lane - pad;
    const bool interior = (x0 >= pad) && (x0 + kFW - pad + 8 < A.cols);
    // raw fetch of one source row: two aligned words (interior) or four clamped bytes packed (edge strips);
    // the funnel shift is applied when the row is consumed, one row later, so the loads stay in flight
    auto fetch = [&](int Y, uint32_t& lo, uint32_t& hi, uint32_t& sh) {
        const int sy = min(max(Y - pad, 0), A.rows - 1);
        const uint8_t* row = src + (size_t)sy * A.src_step;
        if (interior) {
            const uint8_t* p = row + xs0;
            sh = 8u * ((uint32_t)(uintptr_t)p & 3u);
            const uint32_t* a = reinterpret_cast<const uint32_t*>(p - ((uintptr_t)p & 3u));
            lo = __ldg(a); hi = __ldg(a + 1);
        } else {
            uint32_t w = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) w |= (uint32_t)__ldg(row + min(max(xs0 + i, 0), A.cols - 1)) << (8 * i);
            lo = w; hi = 0; sh = 0;
        }
    };

    const int xo = x0 + 4 * lane;                       // output column of byte 0
    const bool lane_out = (4 * lane < A.ow) && (4 * lane + 3 + d < kFW) && xo < A.out_cols;
    const bool full4 = xo + 3 < A.out_cols;
    uint32_t vmask = 0;                                 // valid output pixels of this lane
#pragma unroll
    for (int i = 0; i < 4; ++i) if (lane_out && xo + i < A.out_cols) vmask |= 1u << i;
    const uint32_t ownmask = (4 * lane < A.ow) ? 0xffffffffu : 0u;  // ow is a multiple of 4
    uint8_t* op = A.dst + (size_t)page * A.dst_page_stride + (size_t)yo0 * A.dst_step + xo;   // running output pointer
    const bool st32 = full4 && ((((uintptr_t)op) | A.dst_step) & 3u) == 0;
    uint32_t* bsum = A.blocksum + (((size_t)page * A.nblk) * A.ns + strip) * 2;

    // zero the ring: rows above the band contribute nothing to the running sums
    for (int i = lane; i < RR * 32; i += 32) pixr[i] = 0;
    __syncwarp();

    uint32_t vS[4] = {0, 0, 0, 0}, vQ[4] = {0, 0, 0, 0};   // column sums over the last d rows
    uint32_t accS = 0, accQ = 0;                        // this lane's share of the current 32-row block sum
    bool over = false;                                  // this lane saw the page's fixup list overflow
    uint32_t nlo, nhi, nsh;
    fetch(Ybeg, nlo, nhi, nsh);
    for (int Y = Ybeg; Y <= Yend; ++Y) {
        const uint32_t w = nsh ? __funnelshift_r(nlo, nhi, nsh) : nlo;
        if (Y < Yend) fetch(Y + 1, nlo, nhi, nsh);
        // block sums of the strip's own columns (only the brute-force fixup reads them)
        if (Y >= own0) {
            const uint32_t wm = w & ownmask;
            accS = __dp4a(wm, 0x01010101u, accS);
            accQ = __dp4a(wm, wm, accQ);
            if ((Y & 31) == 31 || Y == Yend) {
                const uint32_t ts = __reduce_add_sync(0xffffffffu, accS), tq = __reduce_add_sync(0xffffffffu, accQ);
                if (lane == 0) {
                    uint32_t* b = bsum + (size_t)(Y >> 5) * A.ns * 2;
                    atomicAdd(b, ts); atomicAdd(b + 1, tq);
                }
                accS = accQ = 0;
            }
        }
        // running column sums: + entering row Y, - row Y-d (this lane's own ring words: no sync needed)
        const uint32_t wold = pixr[((Y - d) & (RR - 1)) * 32 + lane];
        pixr[(Y & (RR - 1)) * 32 + lane] = w;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t pn = byte_of(w, i), po = byte_of(wold, i);
            const uint32_t df = pn - po;
            vS[i] += df;
            vQ[i] += df * (pn + po);
        }
        const int yo = Y - d;                           // output row whose window ends with this row
        if (yo < yo0) continue;                         // (uniform across the warp)
        if ((Y & 31) == 0) {
            // the page already has more undecided pixels than the fixup takes: it will be redone, stop working on it
            const uint32_t c = *reinterpret_cast<volatile uint32_t*>(A.scount + page);
            if (__shfl_sync(0xffffffffu, c, 0) > A.cap) return;
        }
        // row prefix of V over the strip
        const uint32_t a1 = vS[0] + vS[1], a2 = a1 + vS[2], a3 = a2 + vS[3];
        const uint32_t b1 = vQ[0] + vQ[1], b2 = b1 + vQ[2], b3 = b2 + vQ[3];
        const uint32_t es = warp_scan_incl(a3) - a3, eq = warp_scan_incl(b3) - b3;
        const uint32_t pS[4] = {es + vS[0], es + a1, es + a2, es + a3};
        const uint32_t pQ[4] = {eq + vQ[0], eq + b1, eq + b2, eq + b3};
        uint32_t* xs = xch + (Y & 1) * (2 * kFW);
        *reinterpret_cast<uint4*>(xs + 4 * lane) = make_uint4(pS[0], pS[1], pS[2], pS[3]);
        *reinterpret_cast<uint4*>(xs + kFW + 4 * lane) = make_uint4(pQ[0], pQ[1], pQ[2], pQ[3]);
        __syncwarp();                                   // prefix row (and every ring row up to Y) visible to all lanes
        uint8_t* orow = op;
        op += A.dst_step;
        if (!lane_out) continue;
        const uint2 s0 = *reinterpret_cast<const uint2*>(xs + 4 * lane + d), s1 = *reinterpret_cast<const uint2*>(xs + 4 * lane + d + 2);
        const uint2 q0 = *reinterpret_cast<const uint2*>(xs + kFW + 4 * lane + d), q1 = *reinterpret_cast<const uint2*>(xs + kFW + 4 * lane + d + 2);
        // exact window sums: columns x+1 .. x+d of the last d rows
        const uint32_t sw[4] = {s0.x - pS[0], s0.y - pS[1], s1.x - pS[2], s1.y - pS[3]};
        const uint32_t qw[4] = {q0.x - pQ[0], q0.y - pQ[1], q1.x - pQ[2], q1.y - pQ[3]};
        // pixels p(yo, xo..xo+3): padded row yo+pad, strip byte offset 4*lane + pad
        uint32_t p4;
        {
            const uint32_t* pr = pixr + ((yo + pad) & (RR - 1)) * 32;
            const int bo = 4 * lane + pad;
            p4 = pr[bo >> 2];
            if (bo & 3) p4 = __funnelshift_r(p4, pr[(bo >> 2) + 1], 8 * (bo & 3));
        }
        // tier 0: variance-free bounds; sign bits instead of predicates
        uint32_t o4 = 0, und = 0;
        float m[4], pf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            m[i] = (float)sw[i] * F.kwf;
            pf[i] = (float)byte_of(p4, i);
            const int a = __float_as_int(fmaf(m[i], F.pa_hi, Hc - pf[i]));      // < 0  <=>  p - a_hi m > Hc : above every T
            const int b = __float_as_int(fmaf(m[i], -F.pa_lo, pf[i] - Lc));     // < 0  <=>  p - a_lo m < Lc : below every T
            const int c = (int)(qw[i] - F.qfloor_u);                            // < 0  <=>  window too dark for the bounds
            o4 |= (uint32_t)((a & ~c) >> 31) & (0xffu << (8 * i));
            und |= ((uint32_t)~((a | b) & ~c) >> 31) << i;
        }
        if (und) {
            // p == 0 is never above its threshold
#pragma unroll
            for (int i = 0; i < 4; ++i) if (pf[i] == 0.f) und &= ~(1u << i);
        }
        und &= vmask;
        if (METHOD != PRL_FENG && und) {
            // tier 1: FP32 estimate from the exact N = w^2 Q - S^2 (all four pixels, branch-free)
            uint32_t low = 0;                           // pixels whose variance is below the conditioning floor
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float fn;
                if (N32) fn = (float)(F.w2 * qw[i] - sw[i] * sw[i]);
                else {
                    const unsigned long long N = (unsigned long long)F.w2 * qw[i] - (unsigned long long)sw[i] * sw[i];
                    fn = fmaf((float)(unsigned int)(N >> 32), 4294967296.0f, (float)(unsigned int)N);
                }
                const float s = (fn * rsqrtf(fn)) * F.inv_w2f;          // fn == 0 gives NaN -> stays undecided
                float T;
                if (METHOD == PRL_SAUVOLA) T = m[i] * fmaf(s, F.c1, F.c2);
                else if (METHOD == PRL_NIBLACK) T = fmaf(F.c0, s, m[i]);
                else { const float r2 = fmaf(m[i], m[i], s * s); T = fmaf(F.c0, r2 * rsqrtf(r2), m[i]); }
                const float g = (pf[i] - 0.5f) - fmaxf(T, 0.0f);
                const bool ok = fn >= F.n_floor;
                const bool d255 = ok && g > mu, d0 = ok && g < -mu;
                if ((und >> i) & 1u) {
                    if (d255) o4 |= 0xffu << (8 * i);
                    if (d255 || d0) und &= ~(1u << i);
                    if (!ok) low |= 1u << i;
                }
            }
            low &= und;
            if (low) {
                // flat windows (s* < s_floor): the reference's s is NaN (-> T8 = 0 -> p > 0) or anywhere in [0, 0.28];
                // decided only if every one of those outcomes gives 255
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!((low >> i) & 1u)) continue;
                    float Ta, Tb;
                    const float sb = 0.28f;
                    if (METHOD == PRL_SAUVOLA) { Ta = m[i] * F.c2; Tb = m[i] * fmaf(sb, F.c1, F.c2); }
                    else if (METHOD == PRL_NIBLACK) { Ta = m[i]; Tb = fmaf(F.c0, sb, m[i]); }
                    else { Ta = fmaf(F.c0, m[i], m[i]); Tb = fmaf(F.c0, sqrtf(fmaf(m[i], m[i], sb * sb)), m[i]); }
                    const float Tmax = fmaxf(fmaxf(Ta, Tb), 0.0f);
                    if ((pf[i] - 0.5f) - Tmax > mu) { o4 |= 0xffu << (8 * i); und &= ~(1u << i); }
                }
            }
        }
        if (und) {
            // tier 2 (FP64 estimate), then the fixup list
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!((und >> i) & 1u)) continue;
                const int r = A.no_tier2 ? -1 : tier2_decide<METHOD>(sw[i], qw[i], byte_of(p4, i), F.kw_d, F.c0_d, F.c1_d, F.c2_d,
                                                                       F.t2_a, F.t2_b, F.t2_vmin, F.w2, imin);
                if (r == 255) o4 |= 0xffu << (8 * i);
                else if (r < 0) {
                    o4 = (o4 & ~(0xffu << (8 * i))) | ((uint32_t)kSentinel << (8 * i));
                    if (!over) {
                        const uint32_t slot = atomicAdd(A.scount + page, 1u);
                        if (slot < A.cap)
                            A.slist[(size_t)page * kPageCap + slot] = ((uint32_t)yo << 16) | (uint32_t)(xo + i);
                        else over = true;
                    }
                }
            }
        }
        if (st32) {
            *reinterpret_cast<uint32_t*>(orow) = o4;
        } else {
            for (int i = 0; i < 4; ++i) if ((vmask >> i) & 1u) orow[i] = (uint8_t)(o4 >> (8 * i));
        }
    }
}

// ---- tier 3: finish the listed pixels with the reference's FP64 arithmetic on the true int64 taps.
// For every listed pixel a whole warp rebuilds the four taps  S[Yt][X] = sum_{Y' <= Yt, X' <= X} P[Y'][X']  from
//   (1) the block sums of whole 32-row blocks x whole strips to the left/above  (written by the fused kernel),
//   (2) the rows of the last, partial block, columns left of the strip           (brute force, lanes over columns),
//   (3) every row up to Yt, columns from the strip start to X                    (brute force, lanes over rows),
// and lane 0 evaluates exact_t8_from_taps.  Pages whose list overflowed are left to the planes path.
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int kFixThreads = 256;

// One CTA per listed pixel (a single warp per pixel took ~0.4 ms for its ~19 k dependent byte loads -- longer than the
// whole two-kernel path needs for a small batch); the entries of all pages are dealt round-robin over the grid.
template <int METHOD>
__global__ void __launch_bounds__(kFixThreads)
fixup_kernel(const FusedArgs A)
{
    __shared__ unsigned long long red[kFixThreads / 32][12];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned int item = 0;
    for (int page = 0; page < A.n_pages; ++page) {
        const uint32_t cnt = A.scount[page];
        if (cnt == 0 || cnt > A.cap) continue;
        const uint8_t* img = A.src + (size_t)page * A.src_page_stride;
        const uint32_t* bs = A.blocksum + (size_t)page * A.nblk * A.ns * 2;
        const double imin = METHOD == PRL_FENG ? (double)A.imin[page] : 0.0;
        for (uint32_t e = 0; e < cnt; ++e, ++item) {
            if (item % gridDim.x != blockIdx.x) continue;
            const uint32_t ent = A.slist[(size_t)page * kPageCap + e];
            const int y = (int)(ent >> 16), x = (int)(ent & 0xffffu);
            const int b = x / A.ow, xs = b * A.ow;                    // strip whose start is <= x
            const int Yb = y + A.d;                                   // bottom tap row
            const int blk_t = (y + 1) >> 5, blk_b = (Yb + 1) >> 5;    // whole blocks at or above the tap rows
            // (1) + (2): everything left of the strip start
            unsigned long long lts = 0, ltq = 0, lbs = 0, lbq = 0;
            for (int i = tid; i < blk_b * b; i += kFixThreads) {
                const int bk = i / b, bb = i - bk * b;
                const uint32_t s = bs[((size_t)bk * A.ns + bb) * 2], q = bs[((size_t)bk * A.ns + bb) * 2 + 1];
                lbs += s; lbq += q;
                if (bk < blk_t) { lts += s; ltq += q; }
            }
            for (int Yp = 32 * blk_t; Yp <= Yb; ++Yp) {
                if (Yp > y && Yp < 32 * blk_b) { Yp = 32 * blk_b - 1; continue; }      // rows already inside whole blocks of the bottom tap
                const uint8_t* row = img + (size_t)min(max(Yp - A.pad, 0), A.rows - 1) * A.src_step;
                unsigned long long s = 0, q = 0;
                for (int X = tid; X < xs; X += kFixThreads) { const unsigned long long p = row[min(max(X - A.pad, 0), A.cols - 1)]; s += p; q += p * p; }
                if (Yp <= y && Yp >= 32 * blk_t) { lts += s; ltq += q; }
                if (Yp >= 32 * blk_b) { lbs += s; lbq += q; }
            }
            // (3): from the strip start to the tap columns, every row
            unsigned long long ta = 0, tb = 0, tc = 0, td = 0, ua = 0, ub = 0, uc = 0, ud = 0;
            for (int Yp = tid; Yp <= Yb; Yp += kFixThreads) {
                const uint8_t* row = img + (size_t)min(max(Yp - A.pad, 0), A.rows - 1) * A.src_step;
                unsigned long long r1s = 0, r1q = 0;
                for (int X = xs; X <= x; ++X) { const unsigned long long p = row[min(max(X - A.pad, 0), A.cols - 1)]; r1s += p; r1q += p * p; }
                unsigned long long r2s = r1s, r2q = r1q;
                for (int X = x + 1; X <= x + A.d; ++X) { const unsigned long long p = row[min(max(X - A.pad, 0), A.cols - 1)]; r2s += p; r2q += p * p; }
                if (Yp <= y) { ta += r1s; tb += r2s; ua += r1q; ub += r2q; }
                tc += r1s; td += r2s; uc += r1q; ud += r2q;
            }
            unsigned long long v[12] = {lts, ltq, lbs, lbq, ta, tb, tc, td, ua, ub, uc, ud};
#pragma unroll
            for (int k = 0; k < 12; ++k) v[k] = warp_sum_u64(v[k]);
            __syncthreads();                                          // the previous item's readers are done with red[]
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 12; ++k) red[wid][k] = v[k];
            }
            __syncthreads();
            if (tid == 0) {
                unsigned long long t[12];
#pragma unroll
                for (int k = 0; k < 12; ++k) { t[k] = 0; for (int w = 0; w < kFixThreads / 32; ++w) t[k] += red[w][k]; }
                const int t8 = exact_t8_from_taps<METHOD>((long long)(t[4] + t[0]), (long long)(t[5] + t[0]), (long long)(t[6] + t[2]), (long long)(t[7] + t[2]),
                                                          (long long)(t[8] + t[1]), (long long)(t[9] + t[1]), (long long)(t[10] + t[3]), (long long)(t[11] + t[3]),
                                                          A.kw, A.p0, A.p1, A.p2, imin, 0.0);
                const int p = img[(size_t)y * A.src_step + x];
                A.dst[(size_t)page * A.dst_page_stride + (size_t)y * A.dst_step + x] = p > t8 ? 255 : 0;
            }
        }
    }
}

// ---- hand-back, decided on the device: the pages whose list overflowed, in page order, for the indirect two-kernel
// launches that follow (api.cu: fused_hand_back).  One CTA; total[0] accumulates over the life of the context.
__global__ void __launch_bounds__(256)
overflow_list_kernel(const uint32_t* __restrict__ scount, int n_pages, uint32_t cap, int* __restrict__ map, int* __restrict__ count,
                     unsigned long long* __restrict__ total)
{
    __shared__ int part[256];
    const int t = threadIdx.x;
    const int per = (n_pages + 255) / 256, p0 = t * per, p1 = min(p0 + per, n_pages);
    int n = 0;
    for (int p = p0; p < p1; ++p) n += scount[p] > cap ? 1 : 0;
    part[t] = n;
    __syncthreads();
    if (t == 0) {
        int run = 0;
        for (int i = 0; i < 256; ++i) { const int v = part[i]; part[i] = run; run += v; }
        *count = run;
        if (run) atomicAdd(total, (unsigned long long)run);
    }
    __syncthreads();
    int o = part[t];
    for (int p = p0; p < p1; ++p)
        if (scount[p] > cap) map[o++] = p;
}

template <int METHOD, int RR>
void launch_fused_t(prl_cuda_ctx* ctx, const FusedArgs& A, const FastArgs& F)
{
    const int warps = A.ns * A.nb * A.n_pages;
    if (F.n32) local_fused_kernel<METHOD, RR, true><<<(warps + 3) / 4, 128, 0, ctx->stream>>>(A, F);
    else local_fused_kernel<METHOD, RR, false><<<(warps + 3) / 4, 128, 0, ctx->stream>>>(A, F);
}

template <int METHOD>
void launch_fused_m(prl_cuda_ctx* ctx, const FusedArgs& A, const FastArgs& F)
{
    if (A.d + 1 <= 16) launch_fused_t<METHOD, 16>(ctx, A, F);
    else launch_fused_t<METHOD, 32>(ctx, A, F);
}

}  // namespace

// Can this call take the fused path?  (mask output only; Wolf-Jolion needs s_max first -> planes path)
bool prl_fused_eligible(const prl_cuda_ctx* ctx, int method, int n_pages, const prl_geom& g, const double* params)
{
    if (!ctx->use_fused || ctx->force_exact) return false;   // default on; set_option("enable_fused", 0) forces the two-kernel path
    if (method == PRL_WOLFJOLION) return false;
    if (g.Wp > 8192) return false;                            // width limit of the hand-back's generic integral kernel
    if ((g.d & 1) || g.d + 1 > 32 || g.d < 2) return false;
    if (g.rows > 65535 || g.cols > 65535) return false;   // fixup list packs (y, x) into 32 bits
    (void)n_pages;
    FastArgs F;
    return fast_margins(method, params, g, &F);
}

int prl_k_fused(prl_cuda_ctx* ctx, int method, const uint8_t* d_src, int n_pages, const prl_geom& g, size_t src_step,
                size_t src_page_stride, const double* params, uint32_t* d_imin, uint8_t* d_dst, size_t dst_step,
                size_t dst_page_stride, const int** d_redo_map, const int** d_redo_count)
{
    FastArgs F;
    if (!fast_margins(method, params, g, &F)) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "fused path not eligible");
    FusedArgs A;
    A.src = d_src; A.src_step = src_step; A.src_page_stride = src_page_stride;
    A.dst = d_dst; A.dst_step = dst_step; A.dst_page_stride = dst_page_stride;
    A.rows = g.rows; A.cols = g.cols; A.pad = g.h; A.d = g.d; A.Hp = g.Hp; A.Wp = g.Wp;
    A.out_rows = g.out_rows; A.out_cols = g.out_cols;
    A.ow = (kFW - g.d) & ~3;
    A.ns = (g.out_cols + A.ow - 1) / A.ow;
    A.n_pages = n_pages;
    A.kw = 1.0 / (double)(g.w * g.w);
    A.p0 = params[0]; A.p1 = 0; A.p2 = 0;
    if (method == PRL_SAUVOLA) { A.p1 = params[0] * (1.0 / 128.0); A.p2 = 1.0 - params[0]; }
    if (method == PRL_FENG) { A.p1 = 1.0 + (1.0 - params[0]); A.p2 = params[2]; }
    A.imin = d_imin;
    A.cap = (uint32_t)std::min(std::max(ctx->fused_page_cap, 0), kPageCap);

    // row bands: enough warps for ~8 waves of 20 warps per SM, bands a multiple of 32 rows and at least 256 -- or at least
    // 64 when the batch is so small that 256-row bands would leave SMs idle (one A4 page: 286 -> 1210 warps, 0.18 -> 0.07 ms;
    // every band re-scans d rows above its first output row, 22 % extra work at 64 rows)
    {
        const long long want = 8LL * ctx->num_sms * 20;
        int min_rows = 256;
        if ((long long)A.ns * n_pages * ((g.out_rows + 255) / 256) < 16LL * ctx->num_sms) min_rows = 64;
        int nb = (int)std::min<long long>((want + (long long)A.ns * n_pages - 1) / ((long long)A.ns * n_pages), (g.out_rows + min_rows - 1) / min_rows);
        nb = std::max(nb, 1);
        A.band_rows = (((g.out_rows + nb - 1) / nb) + 31) & ~31;
        A.nb = (g.out_rows + A.band_rows - 1) / A.band_rows;
    }
    A.nblk = (g.Hp + 31) / 32;
    A.no_tier2 = ctx->fused_no_tier2 ? 1 : 0;
    // workspace: [scount: n_pages u32][redo count + map: 64 + n_pages i32][slist: n_pages * kPageCap u32][blocksum]
    // (counters and block sums start at zero)
    const size_t cnt_bytes = ((size_t)n_pages * 4 + 255) & ~(size_t)255;
    const size_t map_bytes = (256 + (size_t)n_pages * 4 + 255) & ~(size_t)255;
    const size_t list_bytes = (size_t)n_pages * kPageCap * 4;
    const size_t bsum_bytes = (size_t)n_pages * A.nblk * A.ns * 2 * sizeof(uint32_t);
    int rc = prl_ensure(ctx, &ctx->fused_ws, &ctx->fused_ws_bytes, cnt_bytes + map_bytes + list_bytes + bsum_bytes); if (rc) return rc;
    if (!ctx->d_redo_total) {
        PRL_CUDA_TRY(ctx, cudaMalloc((void**)&ctx->d_redo_total, 256));
        PRL_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_redo_total, 0, 256, ctx->stream));
    }
    A.scount = (uint32_t*)ctx->fused_ws;
    int* redo_count = (int*)((uint8_t*)ctx->fused_ws + cnt_bytes);
    int* redo_map = redo_count + 64;
    A.slist = (uint32_t*)((uint8_t*)ctx->fused_ws + cnt_bytes + map_bytes);
    A.blocksum = (uint32_t*)((uint8_t*)ctx->fused_ws + cnt_bytes + map_bytes + list_bytes);
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(A.scount, 0, cnt_bytes, ctx->stream));
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(A.blocksum, 0, bsum_bytes, ctx->stream));
    if (method == PRL_FENG) {
        PRL_CUDA_TRY(ctx, cudaMemsetAsync(d_imin, 0xff, sizeof(uint32_t) * n_pages, ctx->stream));
        prl_launch_scope ls(ctx, FAM_FUSED_PRE);
        page_min_kernel<<<dim3(64, n_pages), 256, 0, ctx->stream>>>(d_src, src_step, src_page_stride, g.rows, g.cols, d_imin);
    }
    {
        prl_launch_scope ls(ctx, FAM_FUSED);
        switch (method) {
        case PRL_SAUVOLA: launch_fused_m<PRL_SAUVOLA>(ctx, A, F); break;
        case PRL_NIBLACK: launch_fused_m<PRL_NIBLACK>(ctx, A, F); break;
        case PRL_NICK:    launch_fused_m<PRL_NICK>(ctx, A, F); break;
        case PRL_FENG:    launch_fused_m<PRL_FENG>(ctx, A, F); break;
        default: return prl_set_err(ctx, PRL_E_INVALID, "method not served by the fused path");
        }
    }
    {
        prl_launch_scope ls(ctx, FAM_FUSED_FIX);
        const int grid = ctx->num_sms * 8;
        switch (method) {
        case PRL_SAUVOLA: fixup_kernel<PRL_SAUVOLA><<<grid, kFixThreads, 0, ctx->stream>>>(A); break;
        case PRL_NIBLACK: fixup_kernel<PRL_NIBLACK><<<grid, kFixThreads, 0, ctx->stream>>>(A); break;
        case PRL_NICK:    fixup_kernel<PRL_NICK><<<grid, kFixThreads, 0, ctx->stream>>>(A); break;
        default:          fixup_kernel<PRL_FENG><<<grid, kFixThreads, 0, ctx->stream>>>(A); break;
        }
    }
    {
        // which pages overflowed their list: decided and listed on the device, no host read-back (the call stays asynchronous)
        prl_launch_scope ls(ctx, FAM_FUSED_FIX);
        overflow_list_kernel<<<1, 256, 0, ctx->stream>>>(A.scount, n_pages, A.cap, redo_map, redo_count, ctx->d_redo_total);
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    *d_redo_map = redo_map;
    *d_redo_count = redo_count;
    return PRL_OK;
}
