/*
 * mol_oracle.c - plain-C restatement of the reference's eval-mode MoL brute-force scoring.  TEST INFRASTRUCTURE ONLY.
 *
 * A second, independent restatement next to oracle/mol_oracle.py (which issues the reference's own ATen ops): scalar
 * fp32 loops, no BLAS, no vector math library, so every operation and its order is visible here.  Only tests/ (and
 * __graft_entry__.build(), which compiles it) use it; nothing under rails_b200/ links or loads it.  Pinned the same way as
 * the Python oracle: tests/test_c_oracle.py checks it against the outputs of the unmodified reference stored in
 * tests/golden/ (the .npz fixtures; scores agree to ~1e-5: same arithmetic, different summation order than ATen's GEMMs).
 *
 * Reference lines followed (relative to the reference root):
 *   rails/similarities/mol/query_embeddings_fns.py:175-254   query sub-embeddings (+ uid hash embeddings)
 *   rails/similarities/layers.py:19-74                       GeGLU / SwiGLU
 *   rails/similarities/mol/item_embeddings_fns.py:149-183    item sub-embeddings
 *   rails/similarities/mol/similarity_fn.py:148-201          gating, glu_silu branch
 *   rails/similarities/mol/similarity_fn.py:31-46            softmax / eval-mode renorm / weighted sum
 *   rails/similarities/mol/similarity_fn.py:341-413          einsum "bnd,xmd->bxnm", / temperature
 *
 * Build: oracle/Makefile -> oracle/_c/libmol_oracle_c.so (gcc -O2 -fopenmp; no -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float silu_f(float x) { return x / (1.0f + expf(-x)); }                 /* torch.nn.functional.silu */
static float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); } /* F.gelu (erf form) */

/* y[n] = sum_k x[k] * w[n*ldw + k] + b[n]      (F.linear: weight (N, K) row-major) */
static void linear_row(const float* x, const float* w, const float* b, float* y, int N, int K) {
  for (int n = 0; n < N; ++n) {
    float acc = 0.0f;
    const float* wr = w + (size_t)n * K;
    for (int k = 0; k < K; ++k) acc += x[k] * wr[k];
    y[n] = acc + (b ? b[n] : 0.0f);
  }
}

/* v / max(||v||_2, eps) over each group of d values   (query_embeddings_fns.py:244-253, item_embeddings_fns.py:173-182) */
static void l2norm_groups(float* v, int groups, int d, float eps) {
  for (int g = 0; g < groups; ++g) {
    float ss = 0.0f;
    for (int i = 0; i < d; ++i) ss += v[g * d + i] * v[g * d + i];
    float nrm = sqrtf(ss);
    if (nrm < eps) nrm = eps;
    for (int i = 0; i < d; ++i) v[g * d + i] /= nrm;
  }
}

/*
 * Item side (item_embeddings_fns.py:165-182, similarity_fn.py:170-171):
 *   xsub (N, PX, d) = l2norm(items @ x_w^T + x_b);   gi (N, L) = silu(items @ gi_w1^T + gi_b1) @ gi_w2^T
 */
void molc_item_side(int64_t N, int Dx, int PX, int d, int L, int Hgi, const float* items, const float* x_w,
                    const float* x_b, const float* gi_w1, const float* gi_b1, const float* gi_w2, float eps,
                    float* xsub, float* gi) {
#pragma omp parallel
  {
    float* h = (float*)malloc(sizeof(float) * (size_t)Hgi);
#pragma omp for schedule(static)
    for (int64_t x = 0; x < N; ++x) {
      const float* e = items + x * Dx;
      float* xs = xsub + x * PX * d;
      linear_row(e, x_w, x_b, xs, PX * d, Dx);
      l2norm_groups(xs, PX, d, eps);
      linear_row(e, gi_w1, gi_b1, h, Hgi, Dx);
      for (int j = 0; j < Hgi; ++j) h[j] = silu_f(h[j]);
      linear_row(h, gi_w2, NULL, gi + x * L, L, Hgi);
    }
    free(h);
  }
}

/*
 * Query side (query_embeddings_fns.py:191-253, layers.py:36-43 / 67-74, similarity_fn.py:166-169):
 *   pre = q @ q_glu_w + q_glu_b  (q_glu_w is (Dq, 2*Hq): NOT transposed, layers.py:40 / :71)
 *   h = act(pre[:Hq]) * pre[Hq:],  act = gelu (geglu, nonlinearity 0) | silu (swiglu, 1)
 *   proj = h @ q_out_w^T + q_out_b -> (PQ - u, d); uid groups appended: table_i[(user_id mod hash_i) + 1]; l2norm
 *   gq (B, L) = silu(q @ gq_w1^T + gq_b1) @ gq_w2^T
 */
void molc_query_side(int B, int Dq, int PQ, int d, int L, int Hq, int Hgq, int nonlinearity, int u,
                     const int32_t* hash_sizes, const float* const* uid_tables, const int64_t* user_ids,
                     const float* q, const float* q_glu_w, const float* q_glu_b, const float* q_out_w,
                     const float* q_out_b, const float* gq_w1, const float* gq_b1, const float* gq_w2, float eps,
                     float* qsub, float* gq) {
  float* pre = (float*)malloc(sizeof(float) * (size_t)2 * Hq);
  float* h = (float*)malloc(sizeof(float) * (size_t)(Hq > Hgq ? Hq : Hgq));
  for (int b = 0; b < B; ++b) {
    const float* qb = q + (size_t)b * Dq;
    for (int j = 0; j < 2 * Hq; ++j) {
      float acc = 0.0f;
      for (int k = 0; k < Dq; ++k) acc += qb[k] * q_glu_w[(size_t)k * 2 * Hq + j];
      pre[j] = acc + q_glu_b[j];
    }
    for (int j = 0; j < Hq; ++j) h[j] = (nonlinearity == 0 ? gelu_f(pre[j]) : silu_f(pre[j])) * pre[Hq + j];
    float* qs = qsub + (size_t)b * PQ * d;
    linear_row(h, q_out_w, q_out_b, qs, (PQ - u) * d, Hq);
    for (int t = 0; t < u; ++t) {
      int64_t row = user_ids[b] % hash_sizes[t];
      if (row < 0) row += hash_sizes[t]; /* torch's % follows the sign of the divisor */
      row += 1;                          /* :205-207 */
      memcpy(qs + (size_t)(PQ - u + t) * d, uid_tables[t] + (size_t)row * d, sizeof(float) * (size_t)d);
    }
    l2norm_groups(qs, PQ, d, eps);
    linear_row(qb, gq_w1, gq_b1, h, Hgq, Dq);
    for (int j = 0; j < Hgq; ++j) h[j] = silu_f(h[j]);
    linear_row(h, gq_w2, NULL, gq + (size_t)b * L, L, Hgq);
  }
  free(pre);
  free(h);
}

/*
 * scores[b, x] for every (query, item) pair (similarity_fn.py:389-413, :172-179, :42-46):
 *   l[n*PX + m] = <qsub[b,n], xsub[x,m]> / tau
 *   g = gq[b] * gi[x] + qi_w2 @ silu(qi_w1 @ l + qi_b1) + qi_b2 ;  w = g * sigmoid(g)
 *   p = softmax(w) ; if renorm: p /= max(sum p, eps) ; score = sum p * l
 */
void molc_scores(int B, int64_t N, int PQ, int PX, int d, int H, const float* qsub, const float* xsub,
                 const float* gq, const float* gi, const float* qi_w1, const float* qi_b1, const float* qi_w2,
                 const float* qi_b2, float tau, float eps, int renorm, float* scores) {
  const int L = PQ * PX;
#pragma omp parallel
  {
    float* l = (float*)malloc(sizeof(float) * (size_t)L);
    float* hid = (float*)malloc(sizeof(float) * (size_t)H);
    float* w = (float*)malloc(sizeof(float) * (size_t)L);
#pragma omp for schedule(static)
    for (int64_t x = 0; x < N; ++x) {
      const float* xs = xsub + x * PX * d;
      const float* gix = gi + x * L;
      for (int b = 0; b < B; ++b) {
        const float* qs = qsub + (size_t)b * PQ * d;
        for (int n = 0; n < PQ; ++n)
          for (int m = 0; m < PX; ++m) {
            float acc = 0.0f;
            for (int i = 0; i < d; ++i) acc += qs[n * d + i] * xs[m * d + i];
            l[n * PX + m] = acc / tau;
          }
        linear_row(l, qi_w1, qi_b1, hid, H, L);
        for (int j = 0; j < H; ++j) hid[j] = silu_f(hid[j]);
        linear_row(hid, qi_w2, qi_b2, w, L, H);
        float mx = -INFINITY;
        for (int j = 0; j < L; ++j) {
          const float g = gq[(size_t)b * L + j] * gix[j] + w[j];
          w[j] = g * (1.0f / (1.0f + expf(-g)));
          if (w[j] > mx) mx = w[j];
        }
        float den = 0.0f;
        for (int j = 0; j < L; ++j) {
          w[j] = expf(w[j] - mx);
          den += w[j];
        }
        float psum = 0.0f, acc = 0.0f;
        for (int j = 0; j < L; ++j) {
          w[j] /= den; /* softmax */
          psum += w[j];
        }
        if (renorm) {
          if (psum < eps) psum = eps;
          for (int j = 0; j < L; ++j) w[j] /= psum;
        }
        for (int j = 0; j < L; ++j) acc += w[j] * l[j];
        scores[(size_t)b * N + x] = acc;
      }
    }
    free(l);
    free(hid);
    free(w);
  }
}

const char* molc_version(void) { return "mol_oracle.c 1 (plain C restatement, test infrastructure)"; }
