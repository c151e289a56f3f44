/* CPU oracle (TEST INFRASTRUCTURE ONLY) — plain-C restatement of the selection stage of
 * render.py:461-469: jax.random.choice(key, N, (K,), replace=False, p) ==
 * lax.top_k(gumbel + log p, K)[1]  [jax 0.9.0.1, jax/_src/random.py].
 * XLA's CPU TopK orders by value (descending) and breaks ties by the lower index [upstream];
 * a stable insertion of each element into a descending list reproduces exactly that order.
 * Parity unpinned against the reference itself (no JAX in this image); cross-checked against
 * the torch stable-sort restatement in oracle/tensorf_oracle.py.
 *
 * Build: make -C oracle   ->  oracle/_build/liboracle_topk.so
 */
#include <stdint.h>
#include <stdlib.h>

/* g (R,N) row-major; idx (R,K) receives the indices in top-k (descending value) order. */
void oracle_topk_select(const float* g, int R, int N, int K, int32_t* idx) {
  for (int r = 0; r < R; ++r) {
    const float* row = g + (int64_t)r * N;
    int32_t* out = idx + (int64_t)r * K;
    int count = 0;
    for (int s = 0; s < N; ++s) {
      /* position of s in the descending list: after every element that is >= row[s]
       * (earlier index wins ties). */
      int pos = count;
      while (pos > 0 && row[out[pos - 1]] < row[s]) --pos;
      if (pos >= K) continue;
      int last = count < K ? count : K - 1;
      for (int j = last; j > pos; --j) out[j] = out[j - 1];
      out[pos] = s;
      if (count < K) ++count;
    }
  }
}
