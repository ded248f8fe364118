/*
 * oracle/needle_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the alignment half of CRISPResso's hot path: EMBOSS 6.6.0
 * `needle` (affine-gap Needleman-Wunsch/Gotoh, EDNAFULL, end gaps free) as the
 * reference invokes it at CRISPResso/CRISPRessoCORE.py:1791-1806 (amplicon),
 * :1812-1824 (HDR amplicon) and :1911-1936 (reverse-complement rescue).
 *
 * The arithmetic lives in a third-party binary that is NOT under /root/reference
 * (pinned: emboss=6.6.0, environment.yml:19; nucleus/embaln.c
 * embAlignPathCalcWithEndGapPenalties / embAlignWalkNWMatrixUsingCompressedTraceback,
 * ajax/core/ajalign.c srspair writer).  This file restates its published algorithm
 * as specified in SURVEY.md Appendix A (A.1 setup, A.2 initialisation, A.3 fill,
 * A.4 start cell + traceback state machine, A.5 walk) and Appendix B.3 (identity
 * rounding).  Pinning: tests/test_oracle_*.py check it against the micro known-answer
 * vectors of SURVEY App. A.7 and, end to end, against the reference's own golden
 * values tests/crispresso_tests.py:181-195.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (crispresso_b200/) never does.
 *
 * Two independent implementations are provided on purpose:
 *   needle_align_f32  -- float32 matrices m/ix/iy + compass, literal to App. A
 *   needle_align_i32  -- exact integer form (App. A.6), scores multiplied by `scale`
 * and the tests assert they agree on every output.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <pthread.h>

#define DIAG 0
#define LEFT 1
#define DOWN 2

/* EDNAFULL restricted to the alphabet CRISPResso admits (CORE:159: A,T,C,G,N):
 * match 5, mismatch -4, N vs base -2, N vs N -1.  U == T.
 * Returns -100 for a character outside that alphabet (caller reports an error). */
static int base_code(char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    case 'N': case 'n': return 4;
    default: return -1;
    }
}

static int ednafull(int ca, int cb)
{
    if (ca == 4 && cb == 4) return -1;
    if (ca == 4 || cb == 4) return -2;
    return ca == cb ? 5 : -4;
}

int oracle_sub_score(char a, char b)
{
    int ca = base_code(a), cb = base_code(b);
    if (ca < 0 || cb < 0) return -100;
    return ednafull(ca, cb);
}

/* E_FPEQ(a,b,e): (b-e) < a && a < (b+e), e = 1.192e-6 (App. A.4) */
static int fpeq(float a, float b)
{
    const float e = 1.192e-6f;
    return ((b - e) < a) && (a < (b + e));
}

typedef struct {
    int alnlen;      /* alignment columns, end gaps included */
    int ident;       /* columns with identical non-gap residues */
    int gaps;        /* gap columns */
    int start1;      /* start cell row (amplicon index) chosen by A.4 */
    int start2;      /* start cell column (read index) */
    int tenths;      /* identity as printed by "%4.1f", times ten (App. B.3) */
    double score;    /* alignment score (needle's "# Score:") */
} oracle_result;

/* App. B.3: sprintf("%4.1f", (float)100. * (float)ident / (float)len) */
int oracle_identity_tenths(int ident, int len)
{
    float f = (100.0f * (float)ident) / (float)len;
    return (int)rint((double)f * 10.0);
}

/* Emit the alignment (App. A.5) given a direction oracle `compass` (row-major
 * [y*Lb+x]) filled along the path, and the start cell.  Writes the three srspair
 * rows; returns the alignment length. */
static int walk(const char *a, int La, const char *b, int Lb, const uint8_t *compass,
                int start1, int start2, char *ref, char *mark, char *qry, int *ident, int *gaps)
{
    int cap = La + Lb;
    char *ra = (char *)malloc((size_t)cap + 1), *rb = (char *)malloc((size_t)cap + 1);
    int n = 0, x, y;
    for (x = Lb - 1; x > start2; x--) { ra[n] = '-'; rb[n] = b[x]; n++; }
    for (y = La - 1; y > start1; y--) { ra[n] = a[y]; rb[n] = '-'; n++; }
    y = start1; x = start2;
    while (x >= 0 && y >= 0) {
        int d = compass[(size_t)y * Lb + x];
        if (d == DIAG)      { ra[n] = a[y]; rb[n] = b[x]; n++; x--; y--; }
        else if (d == LEFT) { ra[n] = '-';  rb[n] = b[x]; n++; x--; }
        else                { ra[n] = a[y]; rb[n] = '-';  n++; y--; }
    }
    for (; x >= 0; x--) { ra[n] = '-'; rb[n] = b[x]; n++; }
    for (; y >= 0; y--) { ra[n] = a[y]; rb[n] = '-'; n++; }
    int id = 0, g = 0;
    for (int i = 0; i < n; i++) {
        char ca = ra[n - 1 - i], cb = rb[n - 1 - i];
        ref[i] = ca; qry[i] = cb;
        if (ca == '-' || cb == '-') { mark[i] = ' '; g++; }
        else {
            int ka = base_code(ca), kb = base_code(cb);
            /* App. B.2: '|' identical residues, ':' positive score (impossible here), '.' else */
            if (ka == kb) { mark[i] = '|'; id++; }
            else mark[i] = (ednafull(ka, kb) > 0) ? ':' : '.';
        }
    }
    ref[n] = mark[n] = qry[n] = 0;
    *ident = id; *gaps = g;
    free(ra); free(rb);
    return n;
}

/* ------------------------------------------------------------------ float32, literal */
int needle_align_f32(const char *a, int La, const char *b, int Lb, float gapopen, float gapextend,
                     char *ref, char *mark, char *qry, oracle_result *res)
{
    if (La < 1 || Lb < 1) return -2;
    for (int i = 0; i < La; i++) if (base_code(a[i]) < 0) return -1;
    for (int i = 0; i < Lb; i++) if (base_code(b[i]) < 0) return -1;
    const float endgapopen = 0.f, endgapextend = 0.f;   /* endweight=false (App. A.1) */
    size_t n = (size_t)La * Lb;
    float *m = (float *)malloc(n * sizeof(float)), *ix = (float *)malloc(n * sizeof(float)),
          *iy = (float *)malloc(n * sizeof(float));
    uint8_t *compass = (uint8_t *)calloc(n, 1);
#define SUB(y, x) ((float)ednafull(base_code(a[y]), base_code(b[x])))
#define AT(M, y, x) M[(size_t)(y) * Lb + (x)]
    /* A.2 */
    AT(m, 0, 0) = SUB(0, 0);
    AT(ix, 0, 0) = AT(iy, 0, 0) = -endgapopen - gapopen;
    for (int y = 1; y < La; y++) {
        float o = AT(m, y - 1, 0) - gapopen, e = AT(iy, y - 1, 0) - gapextend;
        AT(iy, y, 0) = o >= e ? o : e;
        AT(m, y, 0) = SUB(y, 0) - (endgapopen + (float)(y - 1) * endgapextend);
        AT(ix, y, 0) = -endgapopen - (float)y * endgapextend - gapopen;
    }
    AT(ix, La - 1, 0) += gapopen - endgapopen;
    for (int x = 1; x < Lb; x++) {
        float o = AT(m, 0, x - 1) - gapopen, e = AT(ix, 0, x - 1) - gapextend;
        AT(ix, 0, x) = o >= e ? o : e;
        AT(m, 0, x) = SUB(0, x) - (endgapopen + (float)(x - 1) * endgapextend);
        AT(iy, 0, x) = -endgapopen - (float)x * endgapextend - gapopen;
    }
    AT(iy, 0, Lb - 1) += gapopen - endgapopen;
    /* A.3 */
    for (int x = 1; x < Lb; x++) {
        for (int y = 1; y < La; y++) {
            float pm = AT(m, y - 1, x - 1), pix = AT(ix, y - 1, x - 1), piy = AT(iy, y - 1, x - 1);
            float best = pm; if (pix > best) best = pix; if (piy > best) best = piy;
            AT(m, y, x) = SUB(y, x) + best;
            float o, e;
            if (x == Lb - 1) { o = AT(m, y - 1, x) - endgapopen; e = AT(iy, y - 1, x) - endgapextend; }
            else {
                float t = AT(m, y - 1, x); if (AT(ix, y - 1, x) > t) t = AT(ix, y - 1, x);
                o = t - gapopen; e = AT(iy, y - 1, x) - gapextend;
            }
            AT(iy, y, x) = o >= e ? o : e;
            if (y == La - 1) { o = AT(m, y, x - 1) - endgapopen; e = AT(ix, y, x - 1) - endgapextend; }
            else {
                float t = AT(m, y, x - 1); if (AT(iy, y, x - 1) > t) t = AT(iy, y, x - 1);
                o = t - gapopen; e = AT(ix, y, x - 1) - gapextend;
            }
            AT(ix, y, x) = o >= e ? o : e;
        }
    }
    /* A.4 start */
    float score = -FLT_MAX; int start1 = La - 1, start2 = Lb - 1;
    for (int x = 0; x < Lb; x++) {
        float v[3] = { AT(m, La - 1, x), AT(ix, La - 1, x), AT(iy, La - 1, x) };
        for (int k = 0; k < 3; k++) if (v[k] > score) { score = v[k]; start2 = x; }
    }
    for (int y = 0; y < La; y++) {
        float v[3] = { AT(m, y, Lb - 1), AT(ix, y, Lb - 1), AT(iy, y, Lb - 1) };
        for (int k = 0; k < 3; k++) if (v[k] > score) { score = v[k]; start1 = y; start2 = Lb - 1; }
    }
    /* A.4 traceback */
    {
        int y = start1, x = start2, prev = 0;
        while (x >= 0 && y >= 0) {
            float mp = AT(m, y, x), cx = AT(ix, y, x), cy = AT(iy, y, x);
            float gex = (y == 0 || y == La - 1) ? endgapextend : gapextend;
            float gey = (x == 0 || x == Lb - 1) ? endgapextend : gapextend;
            int dir;
            if (prev == LEFT && fpeq(cx - AT(ix, y, x + 1), gex)) dir = LEFT;
            else if (prev == DOWN && fpeq(cy - AT(iy, y + 1, x), gey)) dir = DOWN;
            else if (mp >= cx && mp >= cy) {
                if (prev == LEFT && mp == cx) dir = LEFT;
                else if (prev == DOWN && mp == cy) dir = DOWN;
                else dir = DIAG;
            }
            else if (cx >= cy) dir = LEFT;
            else dir = DOWN;
            AT(compass, y, x) = (uint8_t)dir;
            if (dir == DIAG) { x--; y--; } else if (dir == LEFT) x--; else y--;
            prev = dir;
        }
    }
    res->alnlen = walk(a, La, b, Lb, compass, start1, start2, ref, mark, qry, &res->ident, &res->gaps);
    res->start1 = start1; res->start2 = start2; res->score = (double)score;
    res->tenths = oracle_identity_tenths(res->ident, res->alnlen);
    free(m); free(ix); free(iy); free(compass);
    return 0;
#undef SUB
}

/* ------------------------------------------------------------------ exact integer form (A.6) */
int needle_align_i32(const char *a, int La, const char *b, int Lb, int gapopen_s, int gapextend_s, int scale,
                     char *ref, char *mark, char *qry, oracle_result *res)
{
    if (La < 1 || Lb < 1) return -2;
    for (int i = 0; i < La; i++) if (base_code(a[i]) < 0) return -1;
    for (int i = 0; i < Lb; i++) if (base_code(b[i]) < 0) return -1;
    size_t n = (size_t)La * Lb;
    int32_t *m = (int32_t *)malloc(n * 4), *ix = (int32_t *)malloc(n * 4), *iy = (int32_t *)malloc(n * 4);
    uint8_t *compass = (uint8_t *)calloc(n, 1);
    const int go = gapopen_s, ge = gapextend_s;
#define SUBI(y, x) (scale * ednafull(base_code(a[y]), base_code(b[x])))
#define MAX2(p, q) ((p) >= (q) ? (p) : (q))
    AT(m, 0, 0) = SUBI(0, 0); AT(ix, 0, 0) = AT(iy, 0, 0) = -go;
    for (int y = 1; y < La; y++) {
        AT(iy, y, 0) = MAX2(AT(m, y - 1, 0) - go, AT(iy, y - 1, 0) - ge);
        AT(m, y, 0) = SUBI(y, 0);
        AT(ix, y, 0) = -go;
    }
    AT(ix, La - 1, 0) += go;
    for (int x = 1; x < Lb; x++) {
        AT(ix, 0, x) = MAX2(AT(m, 0, x - 1) - go, AT(ix, 0, x - 1) - ge);
        AT(m, 0, x) = SUBI(0, x);
        AT(iy, 0, x) = -go;
    }
    AT(iy, 0, Lb - 1) += go;
    for (int x = 1; x < Lb; x++)
        for (int y = 1; y < La; y++) {
            int32_t best = MAX2(AT(m, y - 1, x - 1), MAX2(AT(ix, y - 1, x - 1), AT(iy, y - 1, x - 1)));
            AT(m, y, x) = SUBI(y, x) + best;
            if (x == Lb - 1) AT(iy, y, x) = MAX2(AT(m, y - 1, x), AT(iy, y - 1, x));
            else AT(iy, y, x) = MAX2(MAX2(AT(m, y - 1, x), AT(ix, y - 1, x)) - go, AT(iy, y - 1, x) - ge);
            if (y == La - 1) AT(ix, y, x) = MAX2(AT(m, y, x - 1), AT(ix, y, x - 1));
            else AT(ix, y, x) = MAX2(MAX2(AT(m, y, x - 1), AT(iy, y, x - 1)) - go, AT(ix, y, x - 1) - ge);
        }
    int64_t score = INT64_MIN; int start1 = La - 1, start2 = Lb - 1;
    for (int x = 0; x < Lb; x++) {
        int32_t v[3] = { AT(m, La - 1, x), AT(ix, La - 1, x), AT(iy, La - 1, x) };
        for (int k = 0; k < 3; k++) if (v[k] > score) { score = v[k]; start2 = x; }
    }
    for (int y = 0; y < La; y++) {
        int32_t v[3] = { AT(m, y, Lb - 1), AT(ix, y, Lb - 1), AT(iy, y, Lb - 1) };
        for (int k = 0; k < 3; k++) if (v[k] > score) { score = v[k]; start1 = y; start2 = Lb - 1; }
    }
    {
        int y = start1, x = start2, prev = 0;
        while (x >= 0 && y >= 0) {
            int32_t mp = AT(m, y, x), cx = AT(ix, y, x), cy = AT(iy, y, x);
            int gex = (y == 0 || y == La - 1) ? 0 : ge;
            int gey = (x == 0 || x == Lb - 1) ? 0 : ge;
            int dir;
            if (prev == LEFT && cx - AT(ix, y, x + 1) == gex) dir = LEFT;
            else if (prev == DOWN && cy - AT(iy, y + 1, x) == gey) dir = DOWN;
            else if (mp >= cx && mp >= cy) {
                if (prev == LEFT && mp == cx) dir = LEFT;
                else if (prev == DOWN && mp == cy) dir = DOWN;
                else dir = DIAG;
            }
            else if (cx >= cy) dir = LEFT;
            else dir = DOWN;
            AT(compass, y, x) = (uint8_t)dir;
            if (dir == DIAG) { x--; y--; } else if (dir == LEFT) x--; else y--;
            prev = dir;
        }
    }
    res->alnlen = walk(a, La, b, Lb, compass, start1, start2, ref, mark, qry, &res->ident, &res->gaps);
    res->start1 = start1; res->start2 = start2; res->score = (double)score / (double)scale;
    res->tenths = oracle_identity_tenths(res->ident, res->alnlen);
    free(m); free(ix); free(iy); free(compass);
    return 0;
#undef SUBI
#undef MAX2
#undef AT
}

/* Batch driver used by the tests and by bench.py's cpu_baseline leg.
 * reads: concatenated bytes, offsets[n+1].  Output strings go to fixed slots of
 * `slot` (>= La + maxLb + 1) bytes each; records to res[n].  use_int selects the integer
 * form.  nthreads <= 1 runs serially (how the reference runs needle: one process);
 * nthreads > 1 uses that many pthreads over a shared work counter. */
typedef struct {
    const char *a; int La; const char *reads; const int64_t *offsets; int64_t n;
    double gapopen, gapextend; int use_int;
    char *ref, *mark, *qry; int64_t slot; oracle_result *res;
    int64_t next; int err; pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    int scale = 1;   /* smallest power of two making both penalties integral (App. A.6) */
    while (scale < 8 && (j->gapopen * scale != floor(j->gapopen * scale) || j->gapextend * scale != floor(j->gapextend * scale)))
        scale *= 2;
    int go = (int)lrint(j->gapopen * scale), ge = (int)lrint(j->gapextend * scale);
    for (;;) {
        int64_t lo, hi;
        pthread_mutex_lock(&j->mu);
        lo = j->next; hi = lo + 64; if (hi > j->n) hi = j->n; j->next = hi;
        pthread_mutex_unlock(&j->mu);
        if (lo >= hi) break;
        for (int64_t i = lo; i < hi; i++) {
            const char *b = j->reads + j->offsets[i];
            int Lb = (int)(j->offsets[i + 1] - j->offsets[i]);
            int rc;
            if (j->use_int)
                rc = needle_align_i32(j->a, j->La, b, Lb, go, ge, scale, j->ref + i * j->slot,
                                      j->mark + i * j->slot, j->qry + i * j->slot, &j->res[i]);
            else
                rc = needle_align_f32(j->a, j->La, b, Lb, (float)j->gapopen, (float)j->gapextend,
                                      j->ref + i * j->slot, j->mark + i * j->slot, j->qry + i * j->slot, &j->res[i]);
            if (rc) { pthread_mutex_lock(&j->mu); j->err = rc; pthread_mutex_unlock(&j->mu); }
        }
    }
    return NULL;
}

int needle_align_batch(const char *a, int La, const char *reads, const int64_t *offsets, int64_t n,
                       double gapopen, double gapextend, int use_int, int nthreads,
                       char *ref, char *mark, char *qry, int64_t slot, oracle_result *res)
{
    batch_job j = { a, La, reads, offsets, n, gapopen, gapextend, use_int, ref, mark, qry, slot, res, 0, 0,
                    PTHREAD_MUTEX_INITIALIZER };
    if (nthreads <= 1) { batch_worker(&j); return j.err; }
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &j);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return j.err;
}
