/* Oracle (test infrastructure): max-sum Viterbi decoding, plain C.
 *
 * CPU restatement of reference inaSpeechSegmenter/pyannote_viterbi.py:118-224
 * for the only way segmenter.py calls it (segmenter.py:72,176): no
 * `consecutive`, no `constraint`, no `initial` => the state-duplication
 * helpers (:51-115) are identities and the prior is log(1/K) (:166-167).
 *
 *   V[0,j]  = E[0,j] + log(1/K)                                   (:194)
 *   P[t,j]  = argmax_k ( V[t-1,k] + A[k,j] )   first max on ties   (:207-210)
 *   V[t,j]  = E[t,j] + ( V[t-1,P[t,j]] + A[P[t,j],j] )             (:213)
 *   X[T-1]  = argmax_j V[T-1,j] ; X[t-1] = P[t, X[t]]              (:217-220)
 *
 * All arithmetic in IEEE double, evaluated in exactly that order, so results
 * are bit-identical to numpy's.  NaN handling follows numpy.argmax: the first
 * NaN wins.  Built by oracle/Makefile into oracle/_build/liboracle.so; used by
 * tests and by bench.py's cpu_baseline leg only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define ORACLE_MAXK 8

static int argmax_first(const double *v, int n)
{
    /* numpy.argmax semantics: first maximum; a NaN is larger than anything. */
    int best = 0;
    double bv = v[0];
    if (isnan(bv)) return 0;
    for (int i = 1; i < n; ++i) {
        if (isnan(v[i])) return i;
        if (v[i] > bv) { bv = v[i]; best = i; }
    }
    return best;
}

/* emission: [T,K] row-major doubles; trans: [K,K] (from,to); states: [T] out.
 * scratch for back-pointers is allocated here. returns 0 on success. */
int oracle_viterbi(const double *emission, const double *trans, int64_t T, int K,
                   int32_t *states)
{
    if (K < 1 || K > ORACLE_MAXK || T < 1) return -1;
    uint8_t *bp = (uint8_t *)malloc((size_t)T * (size_t)K);
    if (!bp) return -2;
    double V[ORACLE_MAXK], Vn[ORACLE_MAXK], cand[ORACLE_MAXK];
    /* np.log(np.ones(k) / k) */
    const double prior = log(1.0 / (double)K);
    for (int j = 0; j < K; ++j) { V[j] = emission[j] + prior; bp[j] = (uint8_t)j; }
    for (int64_t t = 1; t < T; ++t) {
        const double *e = emission + t * K;
        for (int j = 0; j < K; ++j) {
            for (int k = 0; k < K; ++k) cand[k] = V[k] + trans[k * K + j];
            int a = argmax_first(cand, K);
            bp[t * K + j] = (uint8_t)a;
            Vn[j] = e[j] + cand[a];
        }
        for (int j = 0; j < K; ++j) V[j] = Vn[j];
    }
    int x = argmax_first(V, K);
    states[T - 1] = x;
    for (int64_t t = T - 1; t >= 1; --t) {
        x = bp[t * K + x];
        states[t - 1] = x;
    }
    free(bp);
    return 0;
}

/* float32 emissions (np.log(r) of a float32 softmax, segmenter.py:176):
 * numpy promotes float32 + float64 to float64 element-wise, i.e. each emission
 * is widened exactly before the add. */
int oracle_viterbi_f32(const float *emission, const double *trans, int64_t T, int K,
                       int32_t *states)
{
    double *e = (double *)malloc((size_t)T * (size_t)K * sizeof(double));
    if (!e) return -2;
    for (int64_t i = 0; i < T * K; ++i) e[i] = (double)emission[i];
    int rc = oracle_viterbi(e, trans, T, K, states);
    free(e);
    return rc;
}
