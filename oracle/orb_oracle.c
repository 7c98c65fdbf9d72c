/*
 * orb_oracle.c -- CPU ORACLE (test infrastructure only; see orb_oracle.h).
 *
 * Plain-C restatement of the reference's ORB front end.  Every function cites the reference
 * lines it follows (paths relative to /root/reference).  Build: gcc -O2 -ffp-contract=off.
 */
#include "orb_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define PATCH_SIZE 31        /* ORBextractor.cpp:22 */
#define HALF_PATCH_SIZE 15   /* ORBextractor.cpp:23 */
#define EDGE_THRESHOLD 19    /* ORBextractor.cpp:24 */

/* rBRIEF pattern, ORBextractor.cpp:100-358 (data; see tools/gen_pattern.py) */
static const int8_t ORB_PATTERN[1024] = {
#include "orb_pattern_31.inc"
};

/* ------------------------------------------------------------------------------------------ */
/* OpenCV scalar helpers                                                                      */
/* ------------------------------------------------------------------------------------------ */

/* cvRound: round half to even (SSE cvtsd2si under the default rounding mode). */
int orbo_cv_round(double v) { return (int)lrint(v); }
static int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static int cv_ceil(double v) { int i = (int)v; return i + (i < v); }
static short sat_short(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

/* cv::fastAtan2 (degrees), OpenCV core/mathfuncs_core: degree-7 odd polynomial.  Called at
 * ORBextractor.cpp:53.  SURVEY.md Appendix A4. */
float orbo_fast_atan2(float y, float x)
{
    static const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    static const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    static const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    static const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)2.2204460492503131e-16);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)2.2204460492503131e-16);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ------------------------------------------------------------------------------------------ */
/* cv::resize, 8UC1, INTER_LINEAR (called at ORBextractor.cpp:1084).  SURVEY.md Appendix A1.   */
/* ------------------------------------------------------------------------------------------ */
static void linear_coeffs(int ssize, int dsize, int clamp_low_frac, int *ofs, short *c0, short *c1)
{
    double scale = 1.0 / ((double)dsize / ssize);
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= s;
        if (s < 0 && clamp_low_frac) { s = 0; f = 0.f; }
        if (s >= ssize - 1 && clamp_low_frac) { s = ssize - 1; f = 0.f; }
        ofs[d] = s;
        c0[d] = sat_short((int)lrintf((1.f - f) * 2048.f));
        c1[d] = sat_short((int)lrintf(f * 2048.f));
    }
}

void orbo_resize_linear_u8(const uint8_t *src, int sw, int sh, size_t spitch,
                           uint8_t *dst, int dw, int dh, size_t dpitch)
{
    int *xofs = (int *)malloc(sizeof(int) * (size_t)(dw + dh));
    int *yofs = xofs + dw;
    short *cx0 = (short *)malloc(sizeof(short) * 2 * (size_t)(dw + dh));
    short *cx1 = cx0 + dw, *cy0 = cx1 + dw, *cy1 = cy0 + dh;
    int *rows[2];
    rows[0] = (int *)malloc(sizeof(int) * 2 * (size_t)dw);
    rows[1] = rows[0] + dw;
    /* x: offsets clamped and fraction zeroed at both ends; y: offset kept, source rows clamped
     * (OpenCV resize.cpp: xofs/ialpha table vs. the invoker's clip(sy + k, 0, ssize.height)). */
    linear_coeffs(sw, dw, 1, xofs, cx0, cx1);
    linear_coeffs(sh, dh, 0, yofs, cy0, cy1);
    int have[2] = { INT_MIN, INT_MIN };
    for (int dy = 0; dy < dh; ++dy) {
        int sy[2];
        for (int k = 0; k < 2; ++k) {
            int r = yofs[dy] + k;
            sy[k] = r < 0 ? 0 : (r >= sh ? sh - 1 : r);
        }
        /* horizontal pass into int32 rows (HResizeLinear<uchar,int,short>), two-slot row cache */
        int *use[2];
        for (int k = 0; k < 2; ++k) {
            int slot = -1;
            for (int s = 0; s < 2; ++s) if (have[s] == sy[k]) slot = s;
            if (slot < 0) {
                slot = (have[0] == sy[1 - k]) ? 1 : 0;
                const uint8_t *S = src + (size_t)sy[k] * spitch;
                int *D = rows[slot];
                for (int dx = 0; dx < dw; ++dx) {
                    int sx = xofs[dx];
                    int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
                    D[dx] = S[sx] * cx0[dx] + S[sx1] * cx1[dx];
                }
                have[slot] = sy[k];
            }
            use[k] = rows[slot];
        }
        /* vertical pass + pack (VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>) */
        uint8_t *D = dst + (size_t)dy * dpitch;
        int b0 = cy0[dy], b1 = cy1[dy];
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (use[0][dx] >> 4)) >> 16) + ((b1 * (use[1][dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(rows[0]);
    free(cx0);
    free(xofs);
}

/* cv::copyMakeBorder(BORDER_REFLECT_101), ORBextractor.cpp:1086-1092.  SURVEY.md A5. */
static int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

void orbo_reflect101_border(const uint8_t *src, int w, int h, size_t spitch, uint8_t *dst, int border, size_t dpitch)
{
    for (int y = -border; y < h + border; ++y) {
        const uint8_t *S = src + (size_t)reflect101(y, h) * spitch;
        uint8_t *D = dst + (size_t)(y + border) * dpitch;
        for (int x = -border; x < w + border; ++x) D[x + border] = S[reflect101(x, w)];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on 8U (ORBextractor.cpp:1046).           */
/* OpenCV >= 3.4.2/4.0 bit-exact fixed-point path: Q8.8 taps.  SURVEY.md Appendix A2.          */
/* ------------------------------------------------------------------------------------------ */
static const int GK[7] = { 18, 34, 48, 56, 48, 34, 18 };

void orbo_gaussian7_s2_u8(const uint8_t *src, int w, int h, size_t spitch, uint8_t *dst, size_t dpitch)
{
    uint16_t *hb = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * (size_t)h);
    for (int y = 0; y < h; ++y) {
        const uint8_t *S = src + (size_t)y * spitch;
        uint16_t *H = hb + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            if (x >= 3 && x + 3 < w) {
                for (int i = 0; i < 7; ++i) acc += GK[i] * S[x + i - 3];
            } else {
                for (int i = 0; i < 7; ++i) acc += GK[i] * S[reflect101(x + i - 3, w)];
            }
            H[x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t *R[7];
        for (int i = 0; i < 7; ++i) R[i] = hb + (size_t)reflect101(y + i - 3, h) * w;
        uint8_t *D = dst + (size_t)y * dpitch;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int i = 0; i < 7; ++i) acc += (uint32_t)GK[i] * R[i][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    free(hb);
}

/* ------------------------------------------------------------------------------------------ */
/* cv::FAST(TYPE_9_16, nonmaxSuppression = true), called per cell at ORBextractor.cpp:766,771. */
/* SURVEY.md Appendix A3 (OpenCV features2d/fast.cpp FAST_t<16>, fast_score.cpp).              */
/* ------------------------------------------------------------------------------------------ */
static const int RING_DX[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
static const int RING_DY[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

static void ring_offsets(size_t pitch, int off[25])
{
    for (int k = 0; k < 16; ++k) off[k] = RING_DX[k] + RING_DY[k] * (int)pitch;
    for (int k = 16; k < 25; ++k) off[k] = off[k - 16];
}

/* is (ptr) a FAST-9 corner at `threshold`?  (strict compares, >= 9 contiguous of 16).  Same
 * early-out structure as OpenCV's FAST_t<16>: every 9-arc contains one pixel of each opposite
 * pair (k, k+8), so a class (darker = 1 / brighter = 2) must survive all eight pair tests. */
static inline int fast9_class(int x, int lo, int hi) { return x < lo ? 1 : (x > hi ? 2 : 0); }
static int fast9_is_corner_slow(const uint8_t *ptr, const int off[25], int threshold);
static inline int fast9_is_corner(const uint8_t *ptr, const int off[25], int threshold)
{
    /* inlined first pair test: rejects flat neighbourhoods without a call */
    int v = ptr[0], a = ptr[off[0]] - v, b = ptr[off[8]] - v;
    if (a >= -threshold && a <= threshold && b >= -threshold && b <= threshold) return 0;
    return fast9_is_corner_slow(ptr, off, threshold);
}
static int fast9_is_corner_slow(const uint8_t *ptr, const int off[25], int threshold)
{
    int v = ptr[0];
    int lo = v - threshold, hi = v + threshold;
    int d = fast9_class(ptr[off[0]], lo, hi) | fast9_class(ptr[off[8]], lo, hi);
    if (d == 0) return 0;
    d &= fast9_class(ptr[off[2]], lo, hi) | fast9_class(ptr[off[10]], lo, hi);
    d &= fast9_class(ptr[off[4]], lo, hi) | fast9_class(ptr[off[12]], lo, hi);
    d &= fast9_class(ptr[off[6]], lo, hi) | fast9_class(ptr[off[14]], lo, hi);
    if (d == 0) return 0;
    d &= fast9_class(ptr[off[1]], lo, hi) | fast9_class(ptr[off[9]], lo, hi);
    d &= fast9_class(ptr[off[3]], lo, hi) | fast9_class(ptr[off[11]], lo, hi);
    d &= fast9_class(ptr[off[5]], lo, hi) | fast9_class(ptr[off[13]], lo, hi);
    d &= fast9_class(ptr[off[7]], lo, hi) | fast9_class(ptr[off[15]], lo, hi);
    if (d & 1) {
        int count = 0;
        for (int k = 0; k < 25; ++k) {
            if (ptr[off[k]] < lo) { if (++count > 8) return 1; } else count = 0;
        }
    }
    if (d & 2) {
        int count = 0;
        for (int k = 0; k < 25; ++k) {
            if (ptr[off[k]] > hi) { if (++count > 8) return 1; } else count = 0;
        }
    }
    return 0;
}

/* cornerScore<16>(ptr, pixel, threshold) */
static int fast9_corner_score(const uint8_t *ptr, const int off[25], int threshold)
{
    int d[25];
    int v = ptr[0];
    for (int k = 0; k < 25; ++k) d[k] = v - ptr[off[k]];
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = d[k + 1] < d[k + 2] ? d[k + 1] : d[k + 2];
        if (d[k + 3] < a) a = d[k + 3];
        if (a <= a0) continue;
        for (int i = 4; i <= 8; ++i) if (d[k + i] < a) a = d[k + i];
        int t = a < d[k] ? a : d[k];
        if (t > a0) a0 = t;
        t = a < d[k + 9] ? a : d[k + 9];
        if (t > a0) a0 = t;
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = d[k + 1] > d[k + 2] ? d[k + 1] : d[k + 2];
        for (int i = 3; i <= 5; ++i) if (d[k + i] > b) b = d[k + i];
        if (b >= b0) continue;
        for (int i = 6; i <= 8; ++i) if (d[k + i] > b) b = d[k + i];
        int t = b > d[k] ? b : d[k];
        if (t < b0) b0 = t;
        t = b > d[k + 9] ? b : d[k + 9];
        if (t < b0) b0 = t;
    }
    return -b0 - 1;
}

void orbo_fast9_score0(const uint8_t *img, int w, int h, size_t pitch, int16_t *score, size_t score_pitch)
{
    int off[25];
    ring_offsets(pitch, off);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int s = -1;
            if (x >= 3 && y >= 3 && x < w - 3 && y < h - 3) {
                const uint8_t *p = img + (size_t)y * pitch + x;
                if (fast9_is_corner(p, off, 0)) s = fast9_corner_score(p, off, 0);
            }
            score[(size_t)y * score_pitch + x] = (int16_t)s;
        }
}

int orbo_fast9_nms(const uint8_t *img, int w, int h, size_t pitch, int threshold, orbo_cand *out, int cap)
{
    if (w < 7 || h < 7) return 0;
    int off[25];
    ring_offsets(pitch, off);
    /* score map with a zero frame: non-corners and everything outside [3,w-3)x[3,h-3) read 0 */
    int *sc = (int *)calloc((size_t)w * (size_t)h, sizeof(int));
    uint8_t *isc = (uint8_t *)calloc((size_t)w * (size_t)h, 1);
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t *row = img + (size_t)y * pitch;
        for (int x = 3; x < w - 3; ++x)
            if (fast9_is_corner(row + x, off, threshold)) {
                sc[(size_t)y * w + x] = fast9_corner_score(row + x, off, threshold);
                isc[(size_t)y * w + x] = 1;
            }
    }
    int n = 0;
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            if (!isc[(size_t)y * w + x]) continue;
            const int *p = sc + (size_t)y * w + x;
            const int s = p[0];
            if (s > p[-1] && s > p[1] && s > p[-w - 1] && s > p[-w] && s > p[-w + 1] &&
                s > p[w - 1] && s > p[w] && s > p[w + 1]) {
                if (n < cap) { out[n].x = (int16_t)x; out[n].y = (int16_t)y; out[n].score = s; }
                ++n;
            }
        }
    free(isc);
    free(sc);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* ExtractorNode::DivideNode + ORBextractor::DistributeOctTree, ORBextractor.cpp:431-718.      */
/* SURVEY.md Appendix B.  std::list emulated with index links; node "address" order replaced  */
/* by the creation sequence number (tie-break rule documented in orb_oracle.h).                */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int x0, y0, x1, y1; /* UL.x, UL.y, UR.x (= BR.x), BL.y (= BR.y) */
    int *keys;          /* candidate indices, stable (candidate) order */
    int nkeys;
    int prev, next;     /* list links */
    int seq;            /* creation order */
} onode;

typedef struct {
    onode *nodes;
    int nnodes, capnodes;
    int head, tail, size;
} olist;

static int olist_new_node(olist *L)
{
    if (L->nnodes == L->capnodes) {
        L->capnodes = L->capnodes ? L->capnodes * 2 : 256;
        L->nodes = (onode *)realloc(L->nodes, sizeof(onode) * (size_t)L->capnodes);
    }
    onode *nd = &L->nodes[L->nnodes];
    memset(nd, 0, sizeof(*nd));
    nd->prev = nd->next = -1;
    nd->seq = L->nnodes;
    return L->nnodes++;
}
static void olist_push_back(olist *L, int id)
{
    onode *nd = &L->nodes[id];
    nd->prev = L->tail; nd->next = -1;
    if (L->tail >= 0) L->nodes[L->tail].next = id; else L->head = id;
    L->tail = id; L->size++;
}
static void olist_push_front(olist *L, int id)
{
    onode *nd = &L->nodes[id];
    nd->next = L->head; nd->prev = -1;
    if (L->head >= 0) L->nodes[L->head].prev = id; else L->tail = id;
    L->head = id; L->size++;
}
static void olist_erase(olist *L, int id)
{
    onode *nd = &L->nodes[id];
    if (nd->prev >= 0) L->nodes[nd->prev].next = nd->next; else L->head = nd->next;
    if (nd->next >= 0) L->nodes[nd->next].prev = nd->prev; else L->tail = nd->prev;
    L->size--;
}

typedef struct { int nkeys, node, seq; } opending;
static int g_tiebreak_sign = 1;
static int pending_cmp(const void *a, const void *b)
{
    const opending *p = (const opending *)a, *q = (const opending *)b;
    if (p->nkeys != q->nkeys) return p->nkeys < q->nkeys ? -1 : 1;
    if (p->seq != q->seq) return (p->seq < q->seq ? -1 : 1) * g_tiebreak_sign;
    return 0;
}

/* DivideNode :431-487 followed by the "add childs if they contain points" block :573-612 /
 * :645-680.  Returns number of >1-key children appended to pend. */
static int divide_and_push(olist *L, int id, const orbo_cand *c, opending *pend, int *npend)
{
    onode par = L->nodes[id];
    int halfX = (int)ceilf((float)(par.x1 - par.x0) / 2);
    int halfY = (int)ceilf((float)(par.y1 - par.y0) / 2);
    int sx = par.x0 + halfX, sy = par.y0 + halfY;
    int bx0[4] = { par.x0, sx, par.x0, sx }, bx1[4] = { sx, par.x1, sx, par.x1 };
    int by0[4] = { par.y0, par.y0, sy, sy }, by1[4] = { sy, sy, par.y1, par.y1 };
    int cnt[4] = { 0, 0, 0, 0 };
    int *buf = (int *)malloc(sizeof(int) * 4 * (size_t)par.nkeys);
    for (int i = 0; i < par.nkeys; ++i) {
        const orbo_cand *k = &c[par.keys[i]];
        int q = ((float)k->x < sx) ? (((float)k->y < sy) ? 0 : 2) : (((float)k->y < sy) ? 1 : 3);
        buf[q * par.nkeys + cnt[q]++] = par.keys[i];
    }
    int added = 0;
    for (int q = 0; q < 4; ++q) {
        if (cnt[q] == 0) continue;
        int cid = olist_new_node(L);
        onode *nd = &L->nodes[cid];
        nd->x0 = bx0[q]; nd->x1 = bx1[q]; nd->y0 = by0[q]; nd->y1 = by1[q];
        nd->nkeys = cnt[q];
        nd->keys = (int *)malloc(sizeof(int) * (size_t)cnt[q]);
        memcpy(nd->keys, buf + q * par.nkeys, sizeof(int) * (size_t)cnt[q]);
        olist_push_front(L, cid);
        if (cnt[q] > 1) {
            pend[*npend].nkeys = cnt[q]; pend[*npend].node = cid; pend[*npend].seq = nd->seq;
            (*npend)++; added++;
        }
    }
    free(buf);
    return added;
}

int orbo_distribute_octree(const orbo_cand *cands, int n, int width, int height, int N, int tiebreak,
                           orbo_cand *out, int cap)
{
    /* :493-495 */
    if (height <= 0) return 0;
    const int nIni = (int)roundf((float)width / (float)height);
    if (nIni <= 0) return 0; /* reference divides by zero here; not reachable for landscape levels */
    const float hX = (float)width / nIni;

    olist L; memset(&L, 0, sizeof(L)); L.head = L.tail = -1;
    int *rootid = (int *)malloc(sizeof(int) * (size_t)nIni);
    for (int i = 0; i < nIni; ++i) { /* :502-513 */
        int id = olist_new_node(&L);
        onode *nd = &L.nodes[id];
        nd->x0 = (int)(hX * (float)i); nd->x1 = (int)(hX * (float)(i + 1));
        nd->y0 = 0; nd->y1 = height;
        nd->keys = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
        olist_push_back(&L, id);
        rootid[i] = id;
    }
    for (int i = 0; i < n; ++i) { /* :516-520 */
        int r = (int)((float)cands[i].x / hX);
        if (r >= nIni) r = nIni - 1; /* unreachable for x < width; guards the oracle only */
        onode *nd = &L.nodes[rootid[r]];
        nd->keys[nd->nkeys++] = i;
    }
    for (int i = 0; i < nIni; ++i) /* :522-535 */
        if (L.nodes[rootid[i]].nkeys == 0) olist_erase(&L, rootid[i]);
    free(rootid);

    size_t pcap = (size_t)(n > 16 ? n : 16) * 4 + 16;
    opending *pend = (opending *)malloc(sizeof(opending) * pcap);
    opending *prev = (opending *)malloc(sizeof(opending) * pcap);
    int finish = 0;
    while (!finish) { /* :545-693 */
        int prevSize = L.size;
        int nToExpand = 0, npend = 0;
        int it = L.head;
        while (it >= 0) {
            if (L.nodes[it].nkeys == 1) { it = L.nodes[it].next; continue; }
            nToExpand += divide_and_push(&L, it, cands, pend, &npend);
            int nx = L.nodes[it].next;
            olist_erase(&L, it);
            it = nx;
        }
        if (L.size >= N || L.size == prevSize) {
            finish = 1;
        } else if (L.size + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = L.size;
                int nprev = npend;
                memcpy(prev, pend, sizeof(opending) * (size_t)nprev);
                npend = 0;
                g_tiebreak_sign = tiebreak ? -1 : 1;
                qsort(prev, (size_t)nprev, sizeof(opending), pending_cmp);
                for (int j = nprev - 1; j >= 0; --j) {
                    divide_and_push(&L, prev[j].node, cands, pend, &npend);
                    olist_erase(&L, prev[j].node);
                    if (L.size >= N) break;
                }
                if (L.size >= N || L.size == prevSize) finish = 1;
            }
        }
    }
    /* :697-715 keep the max-response key of every node (first wins) */
    int nout = 0;
    for (int it = L.head; it >= 0; it = L.nodes[it].next) {
        const onode *nd = &L.nodes[it];
        int best = nd->keys[0];
        for (int k = 1; k < nd->nkeys; ++k)
            if ((float)cands[nd->keys[k]].score > (float)cands[best].score) best = nd->keys[k];
        if (nout < cap) out[nout] = cands[best];
        ++nout;
    }
    for (int i = 0; i < L.nnodes; ++i) free(L.nodes[i].keys);
    free(L.nodes); free(pend); free(prev);
    return nout;
}

/* ------------------------------------------------------------------------------------------ */
/* IC_Angle, ORBextractor.cpp:27-54                                                           */
/* ------------------------------------------------------------------------------------------ */
float orbo_ic_angle(const uint8_t *img, size_t pitch, int x, int y, const int32_t *umax, int32_t *pm10, int32_t *pm01)
{
    int m_01 = 0, m_10 = 0;
    const uint8_t *center = img + (size_t)y * pitch + x;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    int step = (int)pitch;
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    if (pm10) *pm10 = m_10;
    if (pm01) *pm01 = m_01;
    return orbo_fast_atan2((float)m_01, (float)m_10);
}

/* ------------------------------------------------------------------------------------------ */
/* computeOrbDescriptor, ORBextractor.cpp:57-97.  cos/sin resolve to the float overloads in    */
/* the reference translation unit (using namespace std + <cmath>), i.e. cosf/sinf.            */
/* ------------------------------------------------------------------------------------------ */
void orbo_orb_descriptor(const uint8_t *img, size_t pitch, int x, int y, float angle_deg, uint8_t *desc)
{
    const float factorPI = (float)(M_PI / 180.f);
    float angle = angle_deg * factorPI;
    float a = cosf(angle), b = sinf(angle);
    const uint8_t *center = img + (size_t)y * pitch + x;
    const int step = (int)pitch;
    const int8_t *pat = ORB_PATTERN;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            int px0 = pat[4 * k], py0 = pat[4 * k + 1], px1 = pat[4 * k + 2], py1 = pat[4 * k + 3];
            int t0 = center[orbo_cv_round(px0 * b + py0 * a) * step + orbo_cv_round(px0 * a - py0 * b)];
            int t1 = center[orbo_cv_round(px1 * b + py1 * a) * step + orbo_cv_round(px1 * a - py1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Extractor object                                                                           */
/* ------------------------------------------------------------------------------------------ */
struct orbo_extractor {
    int nfeatures, nlevels, iniTh, minTh, tiebreak;
    double scaleFactor; /* ORBextractor.h:84: a double member holding a float value */
    float scale[ORBO_MAX_LEVELS], inv_scale[ORBO_MAX_LEVELS], sigma2[ORBO_MAX_LEVELS], inv_sigma2[ORBO_MAX_LEVELS];
    int nfeat[ORBO_MAX_LEVELS];
    int32_t umax[HALF_PATCH_SIZE + 1];
    /* last run */
    int lw[ORBO_MAX_LEVELS], lh[ORBO_MAX_LEVELS];
    uint8_t *level[ORBO_MAX_LEVELS], *blur[ORBO_MAX_LEVELS];
    size_t levelcap[ORBO_MAX_LEVELS];
    int has_blur[ORBO_MAX_LEVELS];
    orbo_cand *cand[ORBO_MAX_LEVELS]; int ncand[ORBO_MAX_LEVELS], capcand[ORBO_MAX_LEVELS];
    orbo_cand *kept[ORBO_MAX_LEVELS]; int nkept[ORBO_MAX_LEVELS], capkept[ORBO_MAX_LEVELS];
    float *angle[ORBO_MAX_LEVELS];
    int retries[ORBO_MAX_LEVELS];
};

orbo_extractor *orbo_create(int nfeatures, float scaleFactor_, int nlevels, int iniThFAST, int minThFAST)
{
    if (nlevels < 1 || nlevels > ORBO_MAX_LEVELS) return NULL;
    orbo_extractor *ex = (orbo_extractor *)calloc(1, sizeof(*ex));
    ex->nfeatures = nfeatures; ex->nlevels = nlevels; ex->iniTh = iniThFAST; ex->minTh = minThFAST;
    ex->scaleFactor = scaleFactor_;
    /* :365-381 */
    ex->scale[0] = 1.0f; ex->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; ++i) {
        ex->scale[i] = (float)(ex->scale[i - 1] * ex->scaleFactor);
        ex->sigma2[i] = ex->scale[i] * ex->scale[i];
    }
    for (int i = 0; i < nlevels; ++i) {
        ex->inv_scale[i] = 1.0f / ex->scale[i];
        ex->inv_sigma2[i] = 1.0f / ex->sigma2[i];
    }
    /* :385-396 */
    float factor = (float)(1.0f / ex->scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int level = 0; level < nlevels - 1; ++level) {
        ex->nfeat[level] = orbo_cv_round(nDesired);
        sum += ex->nfeat[level];
        nDesired *= factor;
    }
    ex->nfeat[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    /* :404-419 */
    int v, v0, vmax = cv_floor(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
    int vmin = cv_ceil(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) ex->umax[v] = orbo_cv_round(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (ex->umax[v0] == ex->umax[v0 + 1]) ++v0;
        ex->umax[v] = v0;
        ++v0;
    }
    return ex;
}

void orbo_destroy(orbo_extractor *ex)
{
    if (!ex) return;
    for (int l = 0; l < ORBO_MAX_LEVELS; ++l) {
        free(ex->level[l]); free(ex->blur[l]); free(ex->cand[l]); free(ex->kept[l]); free(ex->angle[l]);
    }
    free(ex);
}

void orbo_set_tiebreak(orbo_extractor *ex, int rule) { ex->tiebreak = rule; }

void orbo_tables(const orbo_extractor *ex, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2,
                 int32_t *features_per_level, int32_t *umax)
{
    for (int i = 0; i < ex->nlevels; ++i) {
        if (scale) scale[i] = ex->scale[i];
        if (inv_scale) inv_scale[i] = ex->inv_scale[i];
        if (sigma2) sigma2[i] = ex->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = ex->inv_sigma2[i];
        if (features_per_level) features_per_level[i] = ex->nfeat[i];
    }
    if (umax) for (int i = 0; i <= HALF_PATCH_SIZE; ++i) umax[i] = ex->umax[i];
}

/* ComputePyramid, ORBextractor.cpp:1071-1096 (the 19-px border is never read by the extractor:
 * FAST stays >= 16 px inside, patches reach <= 18 px from keypoints that are >= 19 px inside, and
 * the blur runs on a border-less clone; the oracle therefore keeps tight level images). */
static void compute_pyramid(orbo_extractor *ex, const uint8_t *img, int width, int height, size_t pitch)
{
    for (int level = 0; level < ex->nlevels; ++level) {
        float scale = ex->inv_scale[level];
        int w = orbo_cv_round((float)width * scale), h = orbo_cv_round((float)height * scale);
        if (w < 1) w = 1;
        if (h < 1) h = 1;
        ex->lw[level] = w; ex->lh[level] = h;
        size_t need = (size_t)w * (size_t)h;
        if (ex->levelcap[level] < need) {
            ex->level[level] = (uint8_t *)realloc(ex->level[level], need);
            ex->blur[level] = (uint8_t *)realloc(ex->blur[level], need);
            ex->levelcap[level] = need;
        }
        if (level == 0) {
            for (int y = 0; y < h; ++y) memcpy(ex->level[0] + (size_t)y * w, img + (size_t)y * pitch, (size_t)w);
        } else {
            orbo_resize_linear_u8(ex->level[level - 1], ex->lw[level - 1], ex->lh[level - 1], (size_t)ex->lw[level - 1],
                                  ex->level[level], w, h, (size_t)w);
        }
    }
}

static void push_cand(orbo_extractor *ex, int level, int x, int y, int score)
{
    if (ex->ncand[level] == ex->capcand[level]) {
        ex->capcand[level] = ex->capcand[level] ? ex->capcand[level] * 2 : 4096;
        ex->cand[level] = (orbo_cand *)realloc(ex->cand[level], sizeof(orbo_cand) * (size_t)ex->capcand[level]);
    }
    orbo_cand *c = &ex->cand[level][ex->ncand[level]++];
    c->x = (int16_t)x; c->y = (int16_t)y; c->score = score;
}

/* ComputeKeyPointsOctTree, ORBextractor.cpp:720-811 */
static void compute_keypoints_octree(orbo_extractor *ex)
{
    const float W = 30;
    orbo_cand *cell = NULL; int cellcap = 0;
    for (int level = 0; level < ex->nlevels; ++level) {
        const int cols = ex->lw[level], rows = ex->lh[level];
        const uint8_t *img = ex->level[level];
        const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
        const int maxBorderX = cols - EDGE_THRESHOLD + 3, maxBorderY = rows - EDGE_THRESHOLD + 3;
        ex->ncand[level] = 0; ex->nkept[level] = 0; ex->retries[level] = 0;
        const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols < 1 || nRows < 1 || width <= 0 || height <= 0) continue; /* reference: ceil(x/0) UB; level too small */
        const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
        for (int i = 0; i < nRows; ++i) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; ++j) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                int cw = (int)maxX - (int)iniX, ch = (int)maxY - (int)iniY;
                int need = cw * ch;
                if (need > cellcap) { cellcap = need; cell = (orbo_cand *)realloc(cell, sizeof(orbo_cand) * (size_t)cellcap); }
                const uint8_t *sub = img + (size_t)(int)iniY * cols + (int)iniX;
                int nc = orbo_fast9_nms(sub, cw, ch, (size_t)cols, ex->iniTh, cell, cellcap);
                if (nc == 0) {
                    nc = orbo_fast9_nms(sub, cw, ch, (size_t)cols, ex->minTh, cell, cellcap);
                    ex->retries[level]++;
                }
                for (int k = 0; k < nc; ++k)
                    push_cand(ex, level, cell[k].x + j * wCell, cell[k].y + i * hCell, cell[k].score);
            }
        }
        int N = ex->nfeat[level];
        int cap = ex->ncand[level] + 8;
        if (cap > ex->capkept[level]) {
            ex->capkept[level] = cap;
            ex->kept[level] = (orbo_cand *)realloc(ex->kept[level], sizeof(orbo_cand) * (size_t)cap);
            ex->angle[level] = (float *)realloc(ex->angle[level], sizeof(float) * (size_t)cap);
        }
        ex->nkept[level] = orbo_distribute_octree(ex->cand[level], ex->ncand[level],
                                                  maxBorderX - minBorderX, maxBorderY - minBorderY, N, ex->tiebreak,
                                                  ex->kept[level], cap);
    }
    free(cell);
    /* computeOrientation :809-810 on the un-blurred levels; keypoint = kept + (16,16) */
    for (int level = 0; level < ex->nlevels; ++level)
        for (int k = 0; k < ex->nkept[level]; ++k)
            ex->angle[level][k] = orbo_ic_angle(ex->level[level], (size_t)ex->lw[level],
                                                ex->kept[level][k].x + 16, ex->kept[level][k].y + 16, ex->umax, NULL, NULL);
}

int orbo_extract(orbo_extractor *ex, const uint8_t *img, int width, int height, size_t pitch,
                 orbo_keypoint *kps, uint8_t *desc, int cap)
{
    if (!img || width <= 0 || height <= 0) return 0; /* :1004-1005 */
    compute_pyramid(ex, img, width, height, pitch);
    compute_keypoints_octree(ex);
    int total = 0;
    for (int l = 0; l < ex->nlevels; ++l) total += ex->nkept[l];
    if (kps && total > cap) return -1;
    int offset = 0;
    for (int level = 0; level < ex->nlevels; ++level) { /* :1036-1064 */
        ex->has_blur[level] = 0;
        int n = ex->nkept[level];
        if (n == 0) continue;
        const int w = ex->lw[level], h = ex->lh[level];
        orbo_gaussian7_s2_u8(ex->level[level], w, h, (size_t)w, ex->blur[level], (size_t)w);
        ex->has_blur[level] = 1;
        const int scaledPatchSize = (int)(PATCH_SIZE * ex->scale[level]); /* :794 */
        const float scale = ex->scale[level];
        for (int k = 0; k < n; ++k) {
            int x = ex->kept[level][k].x + 16, y = ex->kept[level][k].y + 16; /* :801-802 */
            float ang = ex->angle[level][k];
            if (desc) orbo_orb_descriptor(ex->blur[level], (size_t)w, x, y, ang, desc + (size_t)(offset + k) * 32);
            if (kps) {
                orbo_keypoint *kp = &kps[offset + k];
                kp->x = (float)x; kp->y = (float)y;
                if (level != 0) { kp->x *= scale; kp->y *= scale; } /* :1055-1061 */
                kp->size = (float)scaledPatchSize;
                kp->angle = ang;
                kp->response = (float)ex->kept[level][k].score;
                kp->octave = level;
                kp->class_id = -1;
            }
        }
        offset += n;
    }
    return total;
}

int orbo_level_dims(const orbo_extractor *ex, int level, int *w, int *h)
{
    if (level < 0 || level >= ex->nlevels) return -1;
    *w = ex->lw[level]; *h = ex->lh[level];
    return 0;
}
const uint8_t *orbo_level_pixels(const orbo_extractor *ex, int level) { return ex->level[level]; }
const uint8_t *orbo_level_blurred(const orbo_extractor *ex, int level) { return ex->has_blur[level] ? ex->blur[level] : NULL; }
int orbo_level_candidates(const orbo_extractor *ex, int level, const orbo_cand **c) { *c = ex->cand[level]; return ex->ncand[level]; }
int orbo_level_kept(const orbo_extractor *ex, int level, const orbo_cand **c) { *c = ex->kept[level]; return ex->nkept[level]; }
int orbo_level_retries(const orbo_extractor *ex, int level) { return ex->retries[level]; }

void orbo_undistort_keypoints(const orbo_keypoint *in, orbo_keypoint *out, int n, const float *cam, const float *dist, int literal_bug)
{
    const double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
    const double k1 = dist[0], k2 = dist[1], p1 = dist[2], p2 = dist[3], k3 = dist[4];
    for (int i = 0; i < n; ++i) {
        out[i] = in[i];
        if (dist[0] == 0.0f) continue;                       /* Frame.cpp:82-86 */
        double x = ((double)in[i].x - cx) / fx, y = ((double)in[i].y - cy) / fy;
        const double x0 = x, y0 = y;
        for (int it = 0; it < 5; ++it) {                     /* cvUndistortPointsInternal, TermCriteria(MAX_ITER, 5) */
            const double r2 = x * x + y * y;
            const double icdist = 1.0 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
            if (icdist < 0) { x = x0; y = y0; break; }
            const double dX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
            const double dY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
            x = (x0 - dX) * icdist; y = (y0 - dY) * icdist;
        }
        const float ux = (float)(x * fx + cx), uy = (float)(y * fy + cy);
        out[i].x = ux; out[i].y = literal_bug ? ux : uy;      /* Frame.cpp:105-106 */
    }
}

void orbo_cvt_gray_u8(const uint8_t *src, int w, int h, size_t spitch, int channels, int rgb_order, uint8_t *dst, size_t dpitch)
{
    for (int y = 0; y < h; ++y) {
        const uint8_t *S = src + (size_t)y * spitch;
        uint8_t *D = dst + (size_t)y * dpitch;
        for (int x = 0; x < w; ++x) {
            const uint8_t *p = S + (size_t)x * channels;
            const int r = rgb_order ? p[0] : p[2], g = p[1], b = rgb_order ? p[2] : p[0];
            D[x] = (uint8_t)((r * 9798 + g * 19235 + b * 3735 + 16384) >> 15);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Matcher                                                                                    */
/* ------------------------------------------------------------------------------------------ */
/* ORBmatcher::DescriptorDistance, ORBmatcher.cpp:128-144 */
int orbo_descriptor_distance(const uint8_t *a, const uint8_t *b)
{
    int32_t pa[8], pb[8];
    memcpy(pa, a, 32); memcpy(pb, b, 32);
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        unsigned int v = (unsigned int)(pa[i] ^ pb[i]);
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24);
    }
    return dist;
}

/* best-2 scan, ORBmatcher.cpp:37-62 (without the one-to-one gate at :49-50) */
void orbo_knn2(const uint8_t *q, int nq, const uint8_t *db, int ndb, int index_base,
               int32_t *d1, int32_t *idx1, int32_t *d2)
{
    for (int i = 0; i < nq; ++i) {
        int best = INT_MAX, best2 = INT_MAX, bidx = -1;
        const uint8_t *qa = q + (size_t)i * 32;
        for (int j = 0; j < ndb; ++j) {
            int dist = orbo_descriptor_distance(qa, db + (size_t)j * 32);
            if (dist < best) { best2 = best; best = dist; bidx = j; }
            else if (dist < best2) best2 = dist;
        }
        d1[i] = best; d2[i] = best2; idx1[i] = bidx >= 0 ? bidx + index_base : -1;
    }
}

typedef struct {
    const uint8_t *q, *db; int q0, q1, ndb, base; int32_t *d1, *idx1, *d2;
} knn_job;
static void *knn_worker(void *p)
{
    knn_job *j = (knn_job *)p;
    orbo_knn2(j->q + (size_t)j->q0 * 32, j->q1 - j->q0, j->db, j->ndb, j->base, j->d1 + j->q0, j->idx1 + j->q0, j->d2 + j->q0);
    return NULL;
}
void orbo_knn2_mt(const uint8_t *q, int nq, const uint8_t *db, int ndb, int index_base,
                  int32_t *d1, int32_t *idx1, int32_t *d2, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nq) nthreads = nq > 0 ? nq : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    knn_job *jobs = (knn_job *)malloc(sizeof(knn_job) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        knn_job jb = { q, db, (int)((long long)nq * t / nthreads), (int)((long long)nq * (t + 1) / nthreads), ndb, index_base, d1, idx1, d2 };
        jobs[t] = jb;
        pthread_create(&th[t], NULL, knn_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* ORBmatcher.cpp:65-67 (TH_LOW gate, float ratio test) */
void orbo_ratio_select(const int32_t *d1, const int32_t *idx1, const int32_t *d2, int nq,
                       int th_low, float ratio, int32_t *match)
{
    for (int i = 0; i < nq; ++i) {
        int m = -1;
        if (idx1[i] >= 0 && d1[i] <= th_low && (float)d1[i] < (float)d2[i] * ratio) m = idx1[i];
        match[i] = m;
    }
}

/* Sequential re-scan semantics applied to shard summaries: feeding (d1, then d2) of each shard in
 * ascending shard order into the :52-61 update reproduces the unsharded (d1, idx1, d2). */
void orbo_merge_shards(const int32_t *d1, const int32_t *idx1, const int32_t *d2, int nshards, int nq,
                       int32_t *od1, int32_t *oidx1, int32_t *od2)
{
    for (int i = 0; i < nq; ++i) {
        int best = INT_MAX, best2 = INT_MAX, bidx = -1;
        for (int s = 0; s < nshards; ++s) {
            size_t k = (size_t)s * nq + i;
            if (idx1[k] < 0) continue;
            int v[2] = { d1[k], d2[k] };
            for (int t = 0; t < 2; ++t) {
                int dist = v[t];
                if (dist < best) { best2 = best; best = dist; if (t == 0) bidx = idx1[k]; }
                else if (dist < best2) best2 = dist;
            }
        }
        od1[i] = best; od2[i] = best2; oidx1[i] = bidx;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Frame grid (Frame.cpp:144-168, 219-271) + SearchForInitialization (ORBmatcher.cpp:9-126)    */
/* ------------------------------------------------------------------------------------------ */
#define FRAME_GRID_ROWS 48 /* Frame.h:11 */
#define FRAME_GRID_COLS 64 /* Frame.h:12 */
#define HISTO_LENGTH 30    /* ORBmatcher.cpp:6 */
#define TH_LOW 50          /* ORBmatcher.cpp:7 */

typedef struct { int *idx; int n, cap; } gcell;

static void compute_three_maxima(const int *hsize, int L, int *ind1, int *ind2, int *ind3) /* :147-188 */
{
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; ++i) {
        const int s = hsize[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; *ind3 = *ind2; *ind2 = i; }
        else if (s > max3) { max3 = s; *ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { *ind2 = -1; *ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { *ind3 = -1; }
}

int orbo_search_for_initialization(const orbo_keypoint *kp1, const uint8_t *desc1, int n1,
                                   const orbo_keypoint *kp2, const uint8_t *desc2, int n2,
                                   float *prev, int32_t *m12, int window, float nnratio, int check_ori,
                                   int width, int height, int literal_bug)
{
    const float minX = 0, maxX = (float)width, minY = 0, maxY = (float)height; /* Frame.cpp:111-119 */
    const float wInv = (float)FRAME_GRID_COLS / (maxX - minX), hInv = (float)FRAME_GRID_ROWS / (maxY - minY);
    gcell *grid = (gcell *)calloc(FRAME_GRID_COLS * FRAME_GRID_ROWS, sizeof(gcell));
    for (int i = 0; i < n2; ++i) { /* AssignFeaturesToGrid + GetGridId, Frame.cpp:144-168 */
        double x = kp2[i].x, y = kp2[i].y;
        int ix = (int)round((x - minX) * wInv);
        int iy = (int)round((y - (literal_bug ? maxY : minY)) * hInv);
        if (ix < 0 || ix >= FRAME_GRID_COLS || iy < 0 || iy >= FRAME_GRID_ROWS) continue;
        gcell *c = &grid[ix * FRAME_GRID_ROWS + iy];
        if (c->n == c->cap) { c->cap = c->cap ? c->cap * 2 : 8; c->idx = (int *)realloc(c->idx, sizeof(int) * (size_t)c->cap); }
        c->idx[c->n++] = i;
    }
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) m12[i] = -1;
    int *rot[HISTO_LENGTH]; int rotn[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; ++i) { rot[i] = (int *)malloc(sizeof(int) * (size_t)(n1 + 1)); rotn[i] = 0; }
    const float factor = HISTO_LENGTH / 360.0f;
    int *matchedDist = (int *)malloc(sizeof(int) * (size_t)(n2 + 1));
    int *m21 = (int *)malloc(sizeof(int) * (size_t)(n2 + 1));
    for (int i = 0; i < n2; ++i) { matchedDist[i] = INT_MAX; m21[i] = -1; }
    int *vind = (int *)malloc(sizeof(int) * (size_t)(n2 + 1));

    for (int i1 = 0; i1 < n1; ++i1) {
        int level1 = kp1[i1].octave;
        if (level1 > 0) continue;
        /* GetFeaturesInArea(x, y, r, minLevel = level1, maxLevel = level1), Frame.cpp:219-271 */
        const float x = prev[2 * i1], y = prev[2 * i1 + 1], r = (float)window;
        int nv = 0;
        do {
            int cx0 = (int)floorf((x - minX - r) * wInv); if (cx0 < 0) cx0 = 0;
            if (cx0 >= FRAME_GRID_COLS) break;
            int cx1 = (int)ceilf((x - minX + r) * wInv); if (cx1 > FRAME_GRID_COLS - 1) cx1 = FRAME_GRID_COLS - 1;
            if (cx1 < 0) break;
            int cy0 = (int)floorf((y - minY - r) * hInv); if (cy0 < 0) cy0 = 0;
            if (cy0 >= FRAME_GRID_ROWS) break;
            int cy1 = (int)ceilf((y - minY + r) * hInv); if (cy1 > FRAME_GRID_ROWS - 1) cy1 = FRAME_GRID_ROWS - 1;
            if (cy1 < 0) break;
            const int minLevel = level1, maxLevel = level1;
            const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
            for (int ix = cx0; ix <= cx1; ++ix)
                for (int iy = cy0; iy <= cy1; ++iy) {
                    const gcell *c = &grid[ix * FRAME_GRID_ROWS + iy];
                    for (int j = 0; j < c->n; ++j) {
                        const orbo_keypoint *k = &kp2[c->idx[j]];
                        if (bCheckLevels) {
                            if (k->octave < minLevel) continue;
                            if (maxLevel >= 0 && k->octave > maxLevel) continue;
                        }
                        const float dx = k->x - x, dy = k->y - y;
                        if (fabsf(dx) < r && fabsf(dy) < r) vind[nv++] = c->idx[j];
                    }
                }
        } while (0);
        if (nv == 0) continue;
        const uint8_t *d1 = desc1 + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int t = 0; t < nv; ++t) {
            int i2 = vind[t];
            int dist = orbo_descriptor_distance(d1, desc2 + (size_t)i2 * 32);
            if (matchedDist[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (m21[bestIdx2] >= 0) { m12[m21[bestIdx2]] = -1; nmatches--; }
                m12[i1] = bestIdx2; m21[bestIdx2] = i1; matchedDist[bestIdx2] = bestDist; nmatches++;
                if (check_ori) {
                    float rotv = kp1[i1].angle - kp2[bestIdx2].angle;
                    if (rotv < 0.0) rotv += 360.0f;
                    int bin = (int)roundf(rotv * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rot[bin][rotn[bin]++] = i1;
                }
            }
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotn, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j = 0; j < rotn[i]; ++j) {
                int idx1 = rot[i][j];
                if (m12[idx1] >= 0) { m12[idx1] = -1; nmatches--; }
            }
        }
    }
    for (int i1 = 0; i1 < n1; ++i1)
        if (m12[i1] >= 0) { prev[2 * i1] = kp2[m12[i1]].x; prev[2 * i1 + 1] = kp2[m12[i1]].y; }
    for (int i = 0; i < FRAME_GRID_COLS * FRAME_GRID_ROWS; ++i) free(grid[i].idx);
    free(grid);
    for (int i = 0; i < HISTO_LENGTH; ++i) free(rot[i]);
    free(matchedDist); free(m21); free(vind);
    return nmatches;
}

/* Windowed search with per-query windows.  The window query is the reference's
 * Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) (Frame.cpp:219-271, level filter :245-258); the loop around it is
 * SearchForInitialization's (ORBmatcher.cpp:9-126) with the gate / acceptance made parameters, which also covers upstream
 * ORB-SLAM2's frame-to-frame SearchByProjection (the reference's own body is empty, include/ORBmatcher.h:24 -- PARITY
 * UNPINNED for gate 1; gate 0 with the SearchForInitialization parameters must equal orbo_search_for_initialization,
 * which is pinned against the reference TU). */
int orbo_search_window(const orbo_keypoint *kp1, const uint8_t *desc1, int n1,
                       const orbo_keypoint *kp2, const uint8_t *desc2, int n2,
                       float *centers, int32_t *m12, const orbo_window_params *P)
{
    const float minX = P->use_bounds ? P->min_x : 0, maxX = P->use_bounds ? P->max_x : (float)P->width;   /* Frame.cpp:111-142 */
    const float minY = P->use_bounds ? P->min_y : 0, maxY = P->use_bounds ? P->max_y : (float)P->height;
    const float wInv = (float)FRAME_GRID_COLS / (maxX - minX), hInv = (float)FRAME_GRID_ROWS / (maxY - minY);
    gcell *grid = (gcell *)calloc(FRAME_GRID_COLS * FRAME_GRID_ROWS, sizeof(gcell));
    for (int i = 0; i < n2; ++i) {
        double x = kp2[i].x, y = kp2[i].y;
        int ix = (int)round((x - minX) * wInv);
        int iy = (int)round((y - (P->literal_gridid_bug ? maxY : minY)) * hInv);
        if (ix < 0 || ix >= FRAME_GRID_COLS || iy < 0 || iy >= FRAME_GRID_ROWS) continue;
        gcell *c = &grid[ix * FRAME_GRID_ROWS + iy];
        if (c->n == c->cap) { c->cap = c->cap ? c->cap * 2 : 8; c->idx = (int *)realloc(c->idx, sizeof(int) * (size_t)c->cap); }
        c->idx[c->n++] = i;
    }
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) m12[i] = -1;
    int *rot[HISTO_LENGTH]; int rotn[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; ++i) { rot[i] = (int *)malloc(sizeof(int) * (size_t)(n1 + 1)); rotn[i] = 0; }
    const float factor = HISTO_LENGTH / 360.0f;
    int *matchedDist = (int *)malloc(sizeof(int) * (size_t)(n2 + 1));
    int *m21 = (int *)malloc(sizeof(int) * (size_t)(n2 + 1));
    for (int i = 0; i < n2; ++i) { matchedDist[i] = INT_MAX; m21[i] = -1; }
    int *vind = (int *)malloc(sizeof(int) * (size_t)(n2 + 1));

    for (int i1 = 0; i1 < n1; ++i1) {
        const int level1 = kp1[i1].octave;
        if (level1 < P->query_level_min || level1 > P->query_level_max) continue;
        const float x = centers[2 * i1], y = centers[2 * i1 + 1];
        if (x != x) continue;                                       /* no projection for this query */
        const float r = P->radius * P->level_scale[level1 & 15];
        const int minLevel = P->level_below < 0 ? 0 : level1 - P->level_below;
        const int maxLevel = P->level_above < 0 ? -1 : level1 + P->level_above;
        int nv = 0;
        do {
            int cx0 = (int)floorf((x - minX - r) * wInv); if (cx0 < 0) cx0 = 0;
            if (cx0 >= FRAME_GRID_COLS) break;
            int cx1 = (int)ceilf((x - minX + r) * wInv); if (cx1 > FRAME_GRID_COLS - 1) cx1 = FRAME_GRID_COLS - 1;
            if (cx1 < 0) break;
            int cy0 = (int)floorf((y - minY - r) * hInv); if (cy0 < 0) cy0 = 0;
            if (cy0 >= FRAME_GRID_ROWS) break;
            int cy1 = (int)ceilf((y - minY + r) * hInv); if (cy1 > FRAME_GRID_ROWS - 1) cy1 = FRAME_GRID_ROWS - 1;
            if (cy1 < 0) break;
            const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
            for (int ix = cx0; ix <= cx1; ++ix)
                for (int iy = cy0; iy <= cy1; ++iy) {
                    const gcell *c = &grid[ix * FRAME_GRID_ROWS + iy];
                    for (int j = 0; j < c->n; ++j) {
                        const orbo_keypoint *k = &kp2[c->idx[j]];
                        if (bCheckLevels) {
                            if (k->octave < minLevel) continue;
                            if (maxLevel >= 0 && k->octave > maxLevel) continue;
                        }
                        const float dx = k->x - x, dy = k->y - y;
                        if (fabsf(dx) < r && fabsf(dy) < r) vind[nv++] = c->idx[j];
                    }
                }
        } while (0);
        if (nv == 0) continue;
        const uint8_t *d1 = desc1 + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int t = 0; t < nv; ++t) {
            const int i2 = vind[t];
            const int dist = orbo_descriptor_distance(d1, desc2 + (size_t)i2 * 32);
            if (P->gate == 0 ? matchedDist[i2] <= dist : m21[i2] >= 0) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestIdx2 < 0 || bestDist > P->th_dist) continue;
        if (P->nnratio > 0.f && !(bestDist < (float)bestDist2 * P->nnratio)) continue;
        if (m21[bestIdx2] >= 0) { m12[m21[bestIdx2]] = -1; nmatches--; }
        m12[i1] = bestIdx2; m21[bestIdx2] = i1; matchedDist[bestIdx2] = bestDist; nmatches++;
        if (P->check_orientation) {
            float rotv = kp1[i1].angle - kp2[bestIdx2].angle;
            if (rotv < 0.0) rotv += 360.0f;
            int bin = (int)roundf(rotv * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            rot[bin][rotn[bin]++] = i1;
        }
    }
    if (P->check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotn, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j = 0; j < rotn[i]; ++j) {
                const int idx1 = rot[i][j];
                if (m12[idx1] >= 0) { m12[idx1] = -1; nmatches--; }
            }
        }
    }
    if (P->update_centers)
        for (int i1 = 0; i1 < n1; ++i1)
            if (m12[i1] >= 0) { centers[2 * i1] = kp2[m12[i1]].x; centers[2 * i1 + 1] = kp2[m12[i1]].y; }
    for (int i = 0; i < FRAME_GRID_COLS * FRAME_GRID_ROWS; ++i) free(grid[i].idx);
    free(grid);
    for (int i = 0; i < HISTO_LENGTH; ++i) free(rot[i]);
    free(matchedDist); free(m21); free(vind);
    return nmatches;
}

/* Group-restricted search = upstream ORB-SLAM2's ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) with the vocabulary node of
 * every keypoint given as a group id (the reference's own body is empty, include/ORBmatcher.h:22 -- PARITY UNPINNED).
 * Upstream walks the two FeatureVectors (std::map, ascending node id) in lock step; for a common node, for every F1 feature
 * of the node in ascending index: scan the node's F2 features in ascending index, skip the already matched ones, keep the
 * best two distances (both start at 256), accept on bestDist1 <= TH_LOW and (float)bestDist1 < ratio * (float)bestDist2,
 * then the rotation histogram as in SearchForInitialization.  group 0xffff = keypoint without a node. */
int orbo_search_groups(const orbo_keypoint *kp1, const uint8_t *desc1, const uint16_t *g1, int n1,
                       const orbo_keypoint *kp2, const uint8_t *desc2, const uint16_t *g2, int n2,
                       int32_t *m12, int th_dist, float nnratio, int check_ori)
{
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) m12[i] = -1;
    int *m21 = (int *)malloc(sizeof(int) * (size_t)(n2 + 1));
    for (int i = 0; i < n2; ++i) m21[i] = -1;
    int *rot[HISTO_LENGTH]; int rotn[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; ++i) { rot[i] = (int *)malloc(sizeof(int) * (size_t)(n1 + 1)); rotn[i] = 0; }
    const float factor = HISTO_LENGTH / 360.0f;
    for (int node = 0; node < 0xffff; ++node) {                        /* ascending node id, like the map iteration */
        for (int i1 = 0; i1 < n1; ++i1) {
            if (g1[i1] != node) continue;
            int best1 = 256, best2 = 256, bestIdx = -1;
            for (int i2 = 0; i2 < n2; ++i2) {
                if (g2[i2] != node) continue;
                if (m21[i2] >= 0) continue;
                const int dist = orbo_descriptor_distance(desc1 + (size_t)i1 * 32, desc2 + (size_t)i2 * 32);
                if (dist < best1) { best2 = best1; best1 = dist; bestIdx = i2; }
                else if (dist < best2) best2 = dist;
            }
            if (bestIdx < 0 || best1 > th_dist) continue;
            if (!((float)best1 < nnratio * (float)best2)) continue;
            m12[i1] = bestIdx; m21[bestIdx] = i1; nmatches++;
            if (check_ori) {
                float rotv = kp1[i1].angle - kp2[bestIdx].angle;
                if (rotv < 0.0) rotv += 360.0f;
                int bin = (int)roundf(rotv * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rot[bin][rotn[bin]++] = i1;
            }
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotn, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j = 0; j < rotn[i]; ++j)
                if (m12[rot[i][j]] >= 0) { m12[rot[i][j]] = -1; nmatches--; }
        }
    }
    for (int i = 0; i < HISTO_LENGTH; ++i) free(rot[i]);
    free(m21);
    return nmatches;
}

/* ------------------------------------------------------------------------------------------ */
/* multi-thread extraction helper for the CPU baseline                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int nfeatures, nlevels, ini, min; float sf;
    const uint8_t *imgs; int w, h; size_t stride; int f0, f1; int32_t *counts;
} ext_job;
static void *ext_worker(void *p)
{
    ext_job *j = (ext_job *)p;
    orbo_extractor *ex = orbo_create(j->nfeatures, j->sf, j->nlevels, j->ini, j->min);
    int cap = j->nfeatures * 2 + 64;
    orbo_keypoint *kps = (orbo_keypoint *)malloc(sizeof(orbo_keypoint) * (size_t)cap);
    uint8_t *desc = (uint8_t *)malloc((size_t)cap * 32);
    for (int f = j->f0; f < j->f1; ++f)
        j->counts[f] = orbo_extract(ex, j->imgs + (size_t)f * j->stride, j->w, j->h, (size_t)j->w, kps, desc, cap);
    free(kps); free(desc);
    orbo_destroy(ex);
    return NULL;
}
int orbo_extract_many(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                      const uint8_t *imgs, int width, int height, size_t frame_stride, int nframes,
                      int nthreads, int32_t *counts)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nframes) nthreads = nframes > 0 ? nframes : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    ext_job *jobs = (ext_job *)malloc(sizeof(ext_job) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        ext_job jb = { nfeatures, nlevels, iniThFAST, minThFAST, scaleFactor, imgs, width, height, frame_stride,
                       (int)((long long)nframes * t / nthreads), (int)((long long)nframes * (t + 1) / nthreads), counts };
        jobs[t] = jb;
        pthread_create(&th[t], NULL, ext_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}
