/* ORACLE (test infrastructure only) -- plain-C restatement of the reference's per-sample loops.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this.
 * Each function cites the /root/reference/src lines it follows.  Built by oracle/build.py with gcc
 * (-O2 -ffp-contract=off: IEEE double, no FMA contraction, so the arithmetic is the interpreter's).
 * Pinned against the unmodified reference run under oracle/refharness.py (tests/test_oracle_vs_reference.py)
 * and the doc-figure known-answer vectors (tests/test_known_answers.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* compress.py:23-70  error_bound(origine, diff, mode, value, ...) on ONE plane.
 * orig/diff: int64 samples, element i at [i*stride] (planes are channel-interleaved: compress.py:319 slices
 * [:, img, :, :, channel] and flatten() walks it row-major, rows wrapping).
 * mode: 0 abs, 1 rel, 2 absrel, 3 pwrel.  diff is updated in place.  scratch: 2*n doubles. */
void tzo_error_bound_plane(const int64_t *orig, int64_t *diff, long n, long stride,
                           int mode, double b0, double b1, double *scratch)
{
    if (b0 == 0.0) return;                                   /* :24 lossless */
    double E = 0.0;
    double *Du = scratch, *Dl = scratch + n;
    if (mode == 1 || mode == 2) {
        if (mode == 2 && b1 == 0.0) return;                  /* :35 */
        int64_t mx = orig[0], mn = orig[0];                  /* :31-32 / :36-37 */
        for (long i = 1; i < n; i++) {
            int64_t v = orig[i * stride];
            if (v > mx) mx = v;
            if (v < mn) mn = v;
        }
        if (mode == 1) {
            E = (double)(mx - mn) * b0;                      /* :33 */
        } else {
            double a = fabs(b0), r = (double)(mx - mn) * b1; /* :38-39 */
            E = (a < r) ? a : r;                             /* :40-43 */
        }
    } else if (mode == 0) {
        E = fabs(b0);                                        /* :29 */
    }
    for (long i = 0; i < n; i++) {
        double d = (double)diff[i * stride];
        double e = (mode == 3) ? (double)orig[i * stride] * b0 : E;   /* :45 pwrel: per-pixel bound */
        Du[i] = d + e;                                       /* :47 */
        Dl[i] = d - e;                                       /* :48 */
    }
    double u = INFINITY, l = -INFINITY;                      /* :55-56 */
    long head = 0;
    for (long i = 0; i < n; i++) {                           /* :58 */
        double mnu = (Du[i] < u) ? Du[i] : u;                /* min((u, Du[i])) */
        double mxl = (Dl[i] > l) ? Dl[i] : l;                /* max((l, Dl[i])) */
        if (mnu - mxl < 0.0) {                               /* :60 */
            double mid = (u + l) / 2;                        /* :61 */
            int64_t q = (int64_t)mid;                        /* float -> int64 slice assignment truncates */
            for (long j = head; j < i; j++) diff[j * stride] = q;
            u = INFINITY; l = -INFINITY;                     /* :62-63 */
            head = i;                                        /* :64 */
        }
        if (Du[i] < u) u = Du[i];                            /* :65 */
        if (l < Dl[i]) l = Dl[i];                            /* :66 */
    }
    if (head < n) {
        double mid = (u + l) / 2;                            /* :67 */
        int64_t q = (int64_t)mid;
        for (long j = head; j < n; j++) diff[j * stride] = q;
    }
}

/* compress.py:73-77  finding_difference (encode): y[0]=x[0], y[i]=x[i-1]-x[i], int16 arithmetic. */
void tzo_delta_encode(const int16_t *x, int16_t *y, long n)
{
    if (n <= 0) return;
    y[0] = x[0];
    for (long i = 1; i < n; i++) y[i] = (int16_t)(x[i - 1] - x[i]);
}

/* decompress.py:22-29  finding_difference (decode): x[0]=y[0], x[i]=x[i-1]-y[i], int16 arithmetic. */
void tzo_delta_decode(const int16_t *y, int16_t *x, long n)
{
    if (n <= 0) return;
    int16_t tmp = y[0];
    x[0] = tmp;
    for (long i = 1; i < n; i++) {
        tmp = (int16_t)(tmp - y[i]);
        x[i] = tmp;
    }
}

/* compress.py:84-90 / decompress.py:31-36  replacing_based_on_frequency: the table passes run one after the
 * other on the running result, exactly as the reference does (so value/index collisions behave the same).
 * dir 0: where(result == table[k], k, result)   (compress)
 * dir 1: where(result == k, table[k], result)   (decompress) */
void tzo_replace(int16_t *arr, long n, const int16_t *table, long T, int dir)
{
    for (long k = 0; k < T; k++) {
        int16_t from = dir ? (int16_t)k : table[k];
        int16_t to = dir ? table[k] : (int16_t)k;
        for (long i = 0; i < n; i++)
            if (arr[i] == from) arr[i] = to;
    }
}

/* compress.py:304-314  residual of one frame: trunc_f32(pred*255) - actual, pred cropped from the padded
 * prediction.  pred: float32 [Hp,Wp,C]; actual: u8 [H,W,C]; out int64 [H,W,C]. */
void tzo_residual_frame(const float *pred, const uint8_t *actual, int64_t *out,
                        long H, long W, long C, long Wp)
{
    for (long y = 0; y < H; y++)
        for (long x = 0; x < W; x++)
            for (long c = 0; c < C; c++) {
                volatile float p255 = pred[(y * Wp + x) * C + c] * 255.0f;   /* :307 float32 product */
                out[(y * W + x) * C + c] = (int64_t)p255 - (int64_t)actual[(y * W + x) * C + c]; /* :310-313 */
            }
}
