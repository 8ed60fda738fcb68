/* TEST INFRASTRUCTURE -- not product code. Plain-C restatement of the integer/indexing part of the reference's
 * sliding_window_predict (/root/reference/utils/eval_utils.py): window enumeration (:54-66) and the fold / average of
 * overlapping windows (:78-95). Compiled by oracle/Makefile (and __graft_entry__.build()) into
 * oracle/_build/libfold_oracle.so and used by tests/ as the bit-exact checker of clipebc_window_origins /
 * clipebc_fold_average. Only tests/, smoke() and bench.py's CPU-baseline legs may load it.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* eval_utils.py:54-55  num_rows = int(np.ceil((H - h) / sh) + 1) */
int oracle_num_windows(int image, int window, int stride) {
  return (int)(ceil((double)(image - window) / (double)stride) + 1.0);
}

/* eval_utils.py:59-66  x_start = i * stride; if x_start + window > image: x_start = image - window */
void oracle_window_origins(int image, int window, int stride, int* origins /* [num_windows] */) {
  int n = oracle_num_windows(image, window, stride);
  for (int i = 0; i < n; ++i) {
    int s = i * stride;
    if (s + window > image) s = image - window;
    origins[i] = s;
  }
}

/* eval_utils.py:78-95. preds: [n_rows*n_cols][gh*gw] with gh = wh / r, gw = ww / r; out: [H/r][W/r].
 * Accumulates in float in window order exactly like the numpy loop, then divides by the coverage count. */
int oracle_fold(const float* preds, int H, int W, int wh, int ww, int sh, int sw, int r, float* out) {
  int nr = oracle_num_windows(H, wh, sh), nc = oracle_num_windows(W, ww, sw);
  int Ho = H / r, Wo = W / r;
  int* ro = (int*)malloc(sizeof(int) * (size_t)nr);
  int* co = (int*)malloc(sizeof(int) * (size_t)nc);
  float* cnt = (float*)calloc((size_t)Ho * Wo, sizeof(float));
  if (!ro || !co || !cnt) return 1;
  oracle_window_origins(H, wh, sh, ro);
  oracle_window_origins(W, ww, sw, co);
  memset(out, 0, sizeof(float) * (size_t)Ho * Wo);
  int idx = 0;
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j, ++idx) {
      int x0 = ro[i] / r, x1 = (ro[i] + wh) / r, y0 = co[j] / r, y1 = (co[j] + ww) / r;
      int gw = y1 - y0;
      const float* p = preds + (size_t)idx * (size_t)(x1 - x0) * (size_t)gw;
      for (int x = x0; x < x1; ++x)
        for (int y = y0; y < y1; ++y) {
          out[(size_t)x * Wo + y] += p[(size_t)(x - x0) * gw + (y - y0)];
          cnt[(size_t)x * Wo + y] += 1.0f;
        }
    }
  for (size_t k = 0; k < (size_t)Ho * Wo; ++k) out[k] /= cnt[k];
  free(ro); free(co); free(cnt);
  return 0;
}
