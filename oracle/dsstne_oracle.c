/*
 * dsstne_oracle.c -- CPU restatement (C99 + OpenMP) of DSSTNE's sparse
 * fully-connected hot path.  TEST INFRASTRUCTURE ONLY; see dsstne_oracle.h.
 *
 * Every function cites the reference kernel whose semantics it follows
 * (E = /root/reference/src/amazon/dsstne/engine).  No code is copied: the
 * reference is CUDA SIMT code organised around warps and shared memory; this
 * file states the same arithmetic as plain loops over rows / columns.
 */
#include "dsstne_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_ESCALEF      1073741824.0f            /* ESCALE = 1<<30, E/GpuTypes.h:65-69 */
#define ORC_ONEOVERESC   (1.0 / 1073741824.0)
#define ORC_MIN_ERROR    1.0e-12f                 /* E/NNTypes.h:46 */
#define ORC_MIN_ACT      0.000001f                /* E/NNTypes.h:47 */
#define ORC_MAX_ACT      0.999999f                /* E/NNTypes.h:48 */
#define ORC_MAX_VALUE    999999999999999.0f       /* E/NNTypes.h:49 */

void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_params_default(orc_params* p)
{
    /* E/NNNetwork.cpp:27-58 (defaults) and E/GpuTypes.cpp:475-498 */
    memset(p, 0, sizeof(*p));
    p->denoising_p = 0.0f;
    p->denoising_q = 1.0f;
    p->deltaBoost_one = 1.0f;
    p->deltaBoost_zero = 1.0f;
    p->SMCE_oneTarget = 0.9f;
    p->SMCE_zeroTarget = 0.1f;
    p->SMCE_oneScale = 1.0f;
    p->SMCE_zeroScale = 1.0f;
}

/* example lookup shared by every kernel on the path:
 *   pos = shuffle ? shuffleIndex[position+b] : position+b   (e.g. E/kernels.cu:670)
 *   Indexed datasets add pos = index[pos]                  (E/kernels.cu:751)      */
static inline uint32_t orc_example(const orc_params* p, const orc_csr* d, uint32_t position, uint32_t b)
{
    uint32_t pos = p->bShuffleIndices ? p->pShuffleIndex[position + b] : position + b;
    if (d->index) pos = d->index[pos];
    return pos;
}

/* analog element -> float.  unsigned char is scaled by 1/256 (E/kernels.cu:923,
 * 1171,1582,2383) and char by 1/128 (E/kernels.cu:1246,1657,1908,2415,2518; the
 * one non-indexed, non-denoised Z kernel at :998 says 1/256 -- a reference
 * inconsistency; the oracle uses 1/128 everywhere, like the dense loader :476). */
static inline float orc_value(const orc_csr* d, uint64_t j)
{
    switch (d->dataType) {
    case ORC_DT_FLOAT:  return ((const float*)d->sparseData)[j];
    case ORC_DT_DOUBLE: return (float)((const double*)d->sparseData)[j];
    case ORC_DT_UINT:   return (float)((const uint32_t*)d->sparseData)[j];
    case ORC_DT_INT:    return (float)((const int32_t*)d->sparseData)[j];
    case ORC_DT_LLINT:  return (float)((const int64_t*)d->sparseData)[j];
    case ORC_DT_ULLINT: return (float)((const uint64_t*)d->sparseData)[j];
    case ORC_DT_UCHAR:  return (float)((const unsigned char*)d->sparseData)[j] * (float)(1.0 / 256.0);
    case ORC_DT_CHAR:   return (float)((const signed char*)d->sparseData)[j] * (float)(1.0 / 128.0);
    default:            return 0.0f;
    }
}

/* ------------------------------------------------------------------ a14 */
void orc_clear_unit(float* unit, const float* bias, uint32_t stride, uint32_t batch)
{   /* E/kernels.cu:60-80: unit[b][s] = bias[s] */
    for (uint32_t b = 0; b < batch; b++)
        memcpy(unit + (size_t)b * stride, bias, stride * sizeof(float));
}

void orc_add_bias(float* unit, const float* bias, uint32_t stride, uint32_t batch)
{   /* E/kernels.cu:564-584: unit[b][s] += bias[s] */
    for (uint32_t b = 0; b < batch; b++)
        for (uint32_t s = 0; s < stride; s++)
            unit[(size_t)b * stride + s] += bias[s];
}

/* ------------------------------------------------------------------ a1-a3 */
void orc_sparse_z(const orc_params* p, const orc_csr* d, uint32_t position, uint32_t batch,
                  uint32_t stride, const float* W, float* Z, float beta, int denoised)
{
    const int analog = d->sparseData != NULL;
#pragma omp parallel for schedule(dynamic, 8)
    for (uint32_t b = 0; b < batch; b++) {
        uint32_t ex = orc_example(p, d, position, b);
        uint64_t start = d->sparseStart[ex], end = d->sparseEnd[ex];
        float w = d->dataWeight ? d->dataWeight[ex] : 1.0f;
        float* z = Z + (size_t)b * stride;
        if (start >= end) continue;          /* kernel's while(start<end) never runs */
        if (!denoised) {
            /* Boolean  E/kernels.cu:662-731: unit = beta*unit; unit += w*W[idx][o]
             * Analog   E/kernels.cu:821-894: unit += W[idx][o]*(w*v)            */
            for (uint32_t o = 0; o < stride; o++) z[o] = (beta == 0.0f) ? 0.0f : beta * z[o];
            for (uint64_t j = start; j < end; j++) {
                const float* wr = W + (size_t)d->sparseIndex[j] * stride;
                float s = analog ? w * orc_value(d, j) : w;
                for (uint32_t o = 0; o < stride; o++) z[o] += s * wr[o];
            }
        } else {
            float wq = p->denoising_q * w;
            if (!analog) {
                /* E/kernels.cu:1317-1389: unit = beta*unit; unit += W (kept nnz);
                 * out = (q*w)*unit -- note the beta*old term is scaled too. */
                for (uint32_t o = 0; o < stride; o++) z[o] = (beta == 0.0f) ? 0.0f : beta * z[o];
                for (uint64_t j = start; j < end; j++) {
                    if (d->denoisingRandom[j] < p->denoising_p) continue;
                    const float* wr = W + (size_t)d->sparseIndex[j] * stride;
                    for (uint32_t o = 0; o < stride; o++) z[o] += wr[o];
                }
                for (uint32_t o = 0; o < stride; o++) z[o] = wq * z[o];
            } else {
                /* E/kernels.cu:1555-1630: value = (q*w)*v staged per nnz;
                 * unit = beta*unit + sum W*value                               */
                for (uint32_t o = 0; o < stride; o++) z[o] = (beta == 0.0f) ? 0.0f : beta * z[o];
                for (uint64_t j = start; j < end; j++) {
                    if (d->denoisingRandom[j] < p->denoising_p) continue;
                    const float* wr = W + (size_t)d->sparseIndex[j] * stride;
                    float s = wq * orc_value(d, j);
                    for (uint32_t o = 0; o < stride; o++) z[o] += wr[o] * s;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ a4 */
uint32_t orc_transposed_capacity(const orc_csr* d, uint32_t examples, uint32_t uniqueExamples,
                                 uint32_t N, uint32_t batch, uint32_t* transposedStart)
{
    /* E/NNTypes.cpp:1427-1492 (CalculateSparseDatapointCounts) then
     * E/NNTypes.cpp:1520-1555 (GenerateSparseTransposedMatrix) */
    uint64_t* count = (uint64_t*)calloc(N, sizeof(uint64_t));      /* _vSparseDatapointCount      */
    uint32_t* maxc  = (uint32_t*)calloc(N, sizeof(uint32_t));      /* _vSparseMaxDatapointCount   */
    uint64_t* multi = (uint64_t*)calloc(N, sizeof(uint64_t));      /* _vSparseMultiDatapointCount */
    uint32_t* vc    = (uint32_t*)calloc(N, sizeof(uint32_t));
    uint32_t* exc   = (uint32_t*)calloc(uniqueExamples, sizeof(uint32_t));
    if (d->index) { for (uint32_t i = 0; i < examples; i++) exc[d->index[i]]++; }
    else          { for (uint32_t i = 0; i < uniqueExamples; i++) exc[i] = 1; }
    for (uint32_t i = 0; i < uniqueExamples; i++) {
        for (uint64_t j = d->sparseStart[i]; j < d->sparseEnd[i]; j++) vc[d->sparseIndex[j]]++;
        for (uint64_t j = d->sparseStart[i]; j < d->sparseEnd[i]; j++) {
            uint32_t x = d->sparseIndex[j];
            if (vc[x] > 0) {
                if (vc[x] > maxc[x]) maxc[x] = vc[x];
                if (vc[x] > 1) multi[x] += exc[i];
                count[x] += (uint64_t)exc[i] * vc[x];
                vc[x] = 0;
            }
        }
    }
    uint32_t offset = 0;
    for (uint32_t i = 0; i < N; i++) {
        transposedStart[i] = offset;
        uint64_t size1 = count[i] < batch ? count[i] : batch;
        if (maxc[i] > 1) {
            uint64_t a = (uint64_t)maxc[i] * batch;
            uint64_t c = batch + (uint64_t)(maxc[i] - 1) * multi[i];
            uint64_t size2 = a < c ? a : c;
            if (size2 > size1) size1 = size2;
        }
        offset += (uint32_t)size1;
        offset = ((offset + 31) >> 5) << 5;
    }
    free(count); free(maxc); free(multi); free(vc); free(exc);
    return offset;
}

/* ------------------------------------------------------------------ a5 */
void orc_sparse_transpose(const orc_params* p, const orc_csr* d, uint32_t position, uint32_t batch,
                          int denoised, uint32_t* transposedEnd, uint32_t* transposedIndex,
                          float* transposedData)
{
    /* E/kernels.cu:1980-2043 (Boolean / Weighted), 2110-2255 (Denoised),
     * 2257-2534 (Analog, AnalogDenoised).  Serial over rows => ascending-row
     * order inside each column (canonical order). */
    const int analog = d->sparseData != NULL;
    for (uint32_t b = 0; b < batch; b++) {
        uint32_t ex = orc_example(p, d, position, b);
        float w = d->dataWeight ? d->dataWeight[ex] : 1.0f;
        /* only the weighted Boolean denoised kernel folds q into the payload
         * (E/kernels.cu:2152); the analog ones store w*v (E/kernels.cu:2344-2354)
         * and q is applied once, by the gradient kernel (E/kernels.cu:2547). */
        if (denoised && !analog && d->dataWeight) w *= p->denoising_q;
        for (uint64_t j = d->sparseStart[ex]; j < d->sparseEnd[ex]; j++) {
            if (denoised && d->denoisingRandom[j] < p->denoising_p) continue;
            uint32_t c = d->sparseIndex[j];
            uint32_t pos = transposedEnd[c]++;
            transposedIndex[pos] = b;
            if (transposedData) transposedData[pos] = analog ? w * orc_value(d, j) : w;
        }
    }
}

/* ------------------------------------------------------------------ a6 */
void orc_sparse_wgrad(const orc_params* p, float alpha, float beta, uint32_t m, uint32_t n,
                      const uint32_t* transposedStart, const uint32_t* transposedEnd,
                      const uint32_t* transposedIndex, const float* transposedData,
                      const float* delta, float* dW)
{
    /* E/kernels.cu:2537-2604 (Boolean), 2614-2684 (Analog).  alpha *= q. */
    alpha *= p->denoising_q;
#pragma omp parallel
    {
        int64_t* sum = (int64_t*)malloc((size_t)n * sizeof(int64_t));
#pragma omp for schedule(dynamic, 64)
        for (uint32_t c = 0; c < m; c++) {
            float* g = dW + (size_t)c * n;
            memset(sum, 0, (size_t)n * sizeof(int64_t));
            for (uint32_t e = transposedStart[c]; e < transposedEnd[c]; e++) {
                const float* dr = delta + (size_t)transposedIndex[e] * n;
                if (transposedData) {
                    float v = transposedData[e];
                    for (uint32_t o = 0; o < n; o++) sum[o] += llrintf(ORC_ESCALEF * v * dr[o]);
                } else {
                    for (uint32_t o = 0; o < n; o++) sum[o] += llrintf(ORC_ESCALEF * dr[o]);
                }
            }
            for (uint32_t o = 0; o < n; o++) {
                float old = (beta == 0.0f) ? 0.0f : beta * g[o];
                float fsum = alpha * (float)((double)sum[o] * ORC_ONEOVERESC);
                g[o] = old + fsum;
            }
        }
        free(sum);
    }
}

/* ------------------------------------------------------------------ a9 */
void orc_activation(int activation, float* data, uint32_t batch, uint32_t stride,
                    float slope, float alpha, float lambda)
{
    size_t size = (size_t)batch * stride;
    switch (activation) {
    case ORC_ACT_SIGMOID:   /* E/kActivation.cu:46-56 */
#pragma omp parallel for
        for (size_t i = 0; i < size; i++) data[i] = 1.0f / (1.0f + expf(-data[i]));
        break;
    case ORC_ACT_TANH:
#pragma omp parallel for
        for (size_t i = 0; i < size; i++) data[i] = tanhf(data[i]);
        break;
    case ORC_ACT_RELU:
#pragma omp parallel for
        for (size_t i = 0; i < size; i++) data[i] = fmaxf(0.0f, data[i]);
        break;
    case ORC_ACT_LRELU:
#pragma omp parallel for
        for (size_t i = 0; i < size; i++) data[i] = fmaxf(data[i], data[i] * slope);
        break;
    case ORC_ACT_ELU:
#pragma omp parallel for
        for (size_t i = 0; i < size; i++) { float x = data[i]; data[i] = (x > 0.0f) ? x : alpha * (expf(x) - 1.0f); }
        break;
    case ORC_ACT_SELU:
#pragma omp parallel for
        for (size_t i = 0; i < size; i++) { float x = data[i]; data[i] = (x > 0.0f) ? lambda * x : lambda * alpha * (expf(x) - 1.0f); }
        break;
    case ORC_ACT_SOFTMAX:   /* E/kActivation.cu:155-229: row max, exp sum, a = min(1, e/sum) */
#pragma omp parallel for
        for (uint32_t b = 0; b < batch; b++) {
            float* r = data + (size_t)b * stride;
            float mx = -9999999999.0f;
            for (uint32_t s = 0; s < stride; s++) mx = fmaxf(mx, r[s]);
            double sum = 0.0;
            for (uint32_t s = 0; s < stride; s++) sum += (double)expf(r[s] - mx);
            float norm = 1.0f / (float)sum;
            for (uint32_t s = 0; s < stride; s++) r[s] = fminf(1.0f, expf(r[s] - mx) * norm);
        }
        break;
    case ORC_ACT_LINEAR:
    default:
        break;
    }
}

/* ------------------------------------------------------------------ a7 */
double orc_sparse_loss(const orc_params* p, const orc_csr* d, int errorFunction, int activation,
                       uint32_t position, uint32_t batch, uint32_t stride, const float* unit,
                       int sparseIgnoreZero)
{
    double total = 0.0;
    const int softmax = (activation == ORC_ACT_SOFTMAX);
    const int analog = d->sparseData != NULL;
#pragma omp parallel for reduction(+ : total) schedule(dynamic, 8)
    for (uint32_t b = 0; b < batch; b++) {
        uint32_t ex = orc_example(p, d, position, b);
        uint64_t start = d->sparseStart[ex], end = d->sparseEnd[ex];
        float wd = d->dataWeight ? d->dataWeight[ex] : 1.0f;
        const float* a_row = unit + (size_t)b * stride;
        double e = 0.0;
        switch (errorFunction) {
        case ORC_ERR_L2: {
            /* E/kLoss.cu:595-666 (Boolean); analog E/kLoss.cu:693-800: target t=v */
            float w = 0.5f * wd;
            if (!sparseIgnoreZero)
                for (uint32_t s = 0; s < stride; s++) e += (double)(w * a_row[s] * a_row[s]);
            for (uint64_t j = start; j < end; j++) {
                float a = a_row[d->sparseIndex[j]];
                float t = analog ? orc_value(d, j) : 1.0f;
                if (sparseIgnoreZero) e += (double)(w * ((a - t) * (a - t)));
                else                  e += (double)(w * ((a - t) * (a - t) - a * a));
            }
            break;
        }
        case ORC_ERR_CROSS_ENTROPY:
            if (softmax) {
                /* E/kLoss.cu:1943-1967: w = weighted ? w_b : 1/nnz_row, NZ only */
                float w = d->dataWeight ? wd : 1.0f / (float)(end - start);
                for (uint64_t j = start; j < end; j++) {
                    float a = a_row[d->sparseIndex[j]];
                    e += (double)(-w * logf(fmaxf(ORC_MIN_ERROR, a)));
                }
            } else {
                /* E/kLoss.cu:1749-1839 */
                if (!sparseIgnoreZero)
                    for (uint32_t s = 0; s < stride; s++)
                        e += (double)(-wd * logf(fmaxf(ORC_MIN_ERROR, 1.0f - a_row[s])));
                for (uint64_t j = start; j < end; j++) {
                    float a = a_row[d->sparseIndex[j]];
                    if (sparseIgnoreZero) e += (double)(-wd * logf(fmaxf(ORC_MIN_ERROR, a)));
                    else e += (double)(wd * (-logf(fmaxf(ORC_MIN_ERROR, a)) + logf(fmaxf(ORC_MIN_ERROR, 1.0f - a))));
                }
            }
            break;
        case ORC_ERR_SMCE:
            if (softmax) {
                /* E/kLoss.cu:2566-2599 launches the NZ kernel with
                 * w = oneScale * (weighted ? w_b : 1/nnz_row) */
                float w = p->SMCE_oneScale * (d->dataWeight ? wd : 1.0f / (float)(end - start));
                for (uint64_t j = start; j < end; j++) {
                    float a = a_row[d->sparseIndex[j]];
                    if (a < p->SMCE_oneTarget) e += (double)(-w * logf(fmaxf(ORC_MIN_ERROR, a)));
                }
            } else {
                /* E/kLoss.cu:2213-2327.  Raw pass weight = zeroScale*w_b (intent;
                 * the reference indexes the shuffle table with the flat element
                 * index at :2224-2226, which only matters for weighted data). */
                if (!sparseIgnoreZero) {
                    float w = p->SMCE_zeroScale * wd;
                    for (uint32_t s = 0; s < stride; s++) {
                        float a = a_row[s];
                        if (a > p->SMCE_zeroTarget) e += (double)(-w * logf(fmaxf(ORC_MIN_ERROR, 1.0f - a)));
                    }
                }
                for (uint64_t j = start; j < end; j++) {
                    float a = a_row[d->sparseIndex[j]];
                    if (sparseIgnoreZero) {
                        float w = p->SMCE_oneScale * wd;
                        if (a < p->SMCE_oneTarget) e += (double)(-w * logf(fmaxf(ORC_MIN_ERROR, a)));
                    } else {
                        if (a > p->SMCE_zeroTarget) e += (double)(wd * p->SMCE_zeroScale * logf(fmaxf(ORC_MIN_ERROR, 1.0f - a)));
                        if (a < p->SMCE_oneTarget)  e += (double)(-wd * p->SMCE_oneScale * logf(fmaxf(ORC_MIN_ERROR, a)));
                    }
                }
            }
            break;
        default:
            break;
        }
        total += e;
    }
    return total;
}

/* ------------------------------------------------------------------ a8 */
static inline float orc_l2_deriv(int activation, float a, float slope, float alpha, float lambda)
{   /* f'(x) expressed through the activation value, as the L2 sparse delta
     * kernels use it (E/kDelta.cu:2193-2482) */
    switch (activation) {
    case ORC_ACT_SIGMOID: return a * (1.0f - a);
    case ORC_ACT_TANH:    return 1.0f - a * a;
    case ORC_ACT_RELU:    return (a > 0.0f) ? 1.0f : 0.0f;
    case ORC_ACT_LRELU:   return (a > 0.0f) ? 1.0f : slope;
    case ORC_ACT_ELU:     return (a > 0.0f) ? 1.0f : (a + alpha);
    case ORC_ACT_SELU:    return (a > 0.0f) ? lambda : lambda * alpha * expf(a);
    default:              return 1.0f;       /* Linear, SoftMax raw */
    }
}

void orc_sparse_output_delta(const orc_params* p, const orc_csr* d, int errorFunction,
                             int activation, uint32_t position, uint32_t batch, uint32_t stride,
                             const float* unit, float* delta, int sparseIgnoreZero,
                             float slope, float alpha, float lambda)
{
    const int analog = d->sparseData != NULL;
#pragma omp parallel for schedule(dynamic, 8)
    for (uint32_t b = 0; b < batch; b++) {
        uint32_t ex = orc_example(p, d, position, b);
        uint64_t start = d->sparseStart[ex], end = d->sparseEnd[ex];
        float wd = d->dataWeight ? d->dataWeight[ex] : 1.0f;
        const float* a_row = unit + (size_t)b * stride;
        float* d_row = delta + (size_t)b * stride;
        if (sparseIgnoreZero) memset(d_row, 0, stride * sizeof(float));   /* cudaMemset, e.g. E/kDelta.cu:2578 */
        switch (errorFunction) {
        case ORC_ERR_L2: {
            /* E/kDelta.cu:2193-2232 (sigmoid, deltaBoost), :2234-2482 (others, no boost),
             * :2484-2521 (softmax: raw w*a, NZ a - (weighted?w_b:1/nnz)) */
            float wz = (activation == ORC_ACT_SIGMOID) ? p->deltaBoost_zero * wd : wd;
            float wo = (activation == ORC_ACT_SIGMOID) ? p->deltaBoost_one * wd : wd;
            if (!sparseIgnoreZero)
                for (uint32_t s = 0; s < stride; s++) {
                    float a = a_row[s];
                    d_row[s] = (activation == ORC_ACT_SOFTMAX) ? wd * a
                             : wz * a * orc_l2_deriv(activation, a, slope, alpha, lambda);
                }
            for (uint64_t j = start; j < end; j++) {
                uint32_t c = d->sparseIndex[j];
                float a = a_row[c];
                float t = analog ? orc_value(d, j) : 1.0f;
                if (activation == ORC_ACT_SOFTMAX) {
                    float w = d->dataWeight ? wd : 1.0f / (float)(end - start);
                    d_row[c] = a - w;
                } else
                    d_row[c] = wo * (a - t) * orc_l2_deriv(activation, a, slope, alpha, lambda);
            }
            break;
        }
        case ORC_ERR_CROSS_ENTROPY:
            if (activation == ORC_ACT_SOFTMAX) {
                /* E/kDelta.cu:6588-6596 reuses the L2 softmax kernels (:2484-2521) */
                if (!sparseIgnoreZero)
                    for (uint32_t s = 0; s < stride; s++) d_row[s] = wd * a_row[s];
                float w = d->dataWeight ? wd : 1.0f / (float)(end - start);
                for (uint64_t j = start; j < end; j++) {
                    uint32_t c = d->sparseIndex[j];
                    d_row[c] = a_row[c] - w;
                }
            } else {
                /* E/kDelta.cu:6533-6571 */
                float wz = p->deltaBoost_zero * wd, wo = p->deltaBoost_one * wd;
                if (!sparseIgnoreZero)
                    for (uint32_t s = 0; s < stride; s++) d_row[s] = wz * a_row[s];
                for (uint64_t j = start; j < end; j++) {
                    uint32_t c = d->sparseIndex[j];
                    d_row[c] = wo * (a_row[c] - 1.0f);
                }
            }
            break;
        case ORC_ERR_SMCE:
            if (activation == ORC_ACT_SOFTMAX) {
                /* E/kDelta.cu:7229-7267: raw zeroScale*a (unweighted), NZ a - oneScale*w */
                if (!sparseIgnoreZero)
                    for (uint32_t s = 0; s < stride; s++) {
                        float a = a_row[s];
                        d_row[s] = (a > p->SMCE_zeroTarget) ? p->SMCE_zeroScale * a : 0.0f;
                    }
                float w = p->SMCE_oneScale * (d->dataWeight ? wd : 1.0f / (float)(end - start));
                for (uint64_t j = start; j < end; j++) {
                    uint32_t c = d->sparseIndex[j];
                    float a = a_row[c];
                    d_row[c] = (a < p->SMCE_oneTarget) ? (a - w) : 0.0f;
                }
            } else {
                /* E/kDelta.cu:7182-7226 */
                float wz = p->SMCE_zeroScale * wd, wo = p->SMCE_oneScale * wd;
                if (!sparseIgnoreZero)
                    for (uint32_t s = 0; s < stride; s++) {
                        float a = a_row[s];
                        d_row[s] = (a > p->SMCE_zeroTarget) ? wz * a : 0.0f;
                    }
                for (uint64_t j = start; j < end; j++) {
                    uint32_t c = d->sparseIndex[j];
                    float a = a_row[c];
                    d_row[c] = (a < p->SMCE_oneTarget) ? wo * (a - 1.0f) : 0.0f;
                }
            }
            break;
        default:
            break;
        }
    }
}

/* ------------------------------------------------------------------ a10 */
void orc_sparseness_penalty(uint32_t batch, uint32_t stride, const float* unit, float* delta,
                            float p, float beta)
{   /* E/kDelta.cu:8979-9007 */
#pragma omp parallel for
    for (uint32_t c = 0; c < stride; c++) {
        float pi = 0.0f;
        for (uint32_t b = 0; b < batch; b++) pi += unit[(size_t)b * stride + c];
        pi /= (float)batch;
        pi = fmaxf(ORC_MIN_ACT, fminf(ORC_MAX_ACT, pi));
        float penalty = beta * (-p / pi + (1.0f - p) / (1.0f - pi));
        for (uint32_t b = 0; b < batch; b++) delta[(size_t)b * stride + c] += penalty;
    }
}

void orc_hadamard(int activation, uint64_t size, float scale, const float* unit, float* delta,
                  float slope, float alpha, float lambda)
{   /* E/kDelta.cu:9021-9151 */
    float oneOverScale = 1.0f / scale;
#pragma omp parallel for
    for (uint64_t i = 0; i < size; i++) {
        float x = unit[i], dl = delta[i];
        switch (activation) {
        case ORC_ACT_SIGMOID: dl = x * (1.0f - x) * dl; break;           /* scale ignored, :9023-9031 */
        case ORC_ACT_TANH:    x *= oneOverScale; dl = scale * (1.0f - x * x) * dl; break;
        case ORC_ACT_RELU:    if (x <= 0.0f) dl = 0.0f; break;
        case ORC_ACT_LRELU:   if (x <= 0.0f) dl *= slope; break;
        case ORC_ACT_ELU:     if (x <= 0.0f) dl *= (x + alpha); break;
        case ORC_ACT_SELU:    if (x > 0.0f) dl *= lambda; else dl *= (x + lambda * alpha); break;
        default: break;                                                  /* Linear: no-op */
        }
        delta[i] = dl;
    }
}

/* ------------------------------------------------------------------ a11 */
void orc_gemm_fwd(uint32_t B, uint32_t k, uint32_t n, const float* A, const float* W, float beta, float* C)
{   /* E/NNLayer.cpp:1072-1086: X(L+1) += X(L)*W, row-major */
#pragma omp parallel for schedule(static)
    for (uint32_t b = 0; b < B; b++) {
        float* c = C + (size_t)b * n;
        if (beta == 0.0f) memset(c, 0, n * sizeof(float));
        else if (beta != 1.0f) for (uint32_t j = 0; j < n; j++) c[j] *= beta;
        for (uint32_t i = 0; i < k; i++) {
            float a = A[(size_t)b * k + i];
            const float* w = W + (size_t)i * n;
            for (uint32_t j = 0; j < n; j++) c[j] += a * w[j];
        }
    }
}

void orc_gemm_dw(uint32_t B, uint32_t k, uint32_t n, float alpha, const float* A, const float* D, float beta, float* G)
{   /* E/NNLayer.cpp:2201-2236: dW = alpha * X(L)^T * Delta(L+1) + beta*dW */
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < k; i++) {
        float* g = G + (size_t)i * n;
        float* acc = (float*)calloc(n, sizeof(float));
        for (uint32_t b = 0; b < B; b++) {
            float a = A[(size_t)b * k + i];
            const float* dr = D + (size_t)b * n;
            for (uint32_t j = 0; j < n; j++) acc[j] += a * dr[j];
        }
        for (uint32_t j = 0; j < n; j++) g[j] = ((beta == 0.0f) ? 0.0f : beta * g[j]) + alpha * acc[j];
        free(acc);
    }
}

void orc_gemm_dx(uint32_t B, uint32_t k, uint32_t n, const float* D, const float* W, float beta, float* Dp)
{   /* E/NNLayer.cpp:2254-2287: Delta(L) = Delta(L+1) * W^T + beta*Delta(L) */
#pragma omp parallel for schedule(static)
    for (uint32_t b = 0; b < B; b++) {
        const float* dr = D + (size_t)b * n;
        for (uint32_t i = 0; i < k; i++) {
            const float* w = W + (size_t)i * n;
            float s = 0.0f;
            for (uint32_t j = 0; j < n; j++) s += dr[j] * w[j];
            float* o = Dp + (size_t)b * k + i;
            *o = ((beta == 0.0f) ? 0.0f : beta * *o) + s;
        }
    }
}

/* ------------------------------------------------------------------ a12 */
static inline float orc_sgn(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }   /* E/kernels.h:19 */

void orc_update_weights(int mode, float alpha, float lambda, float lambda1, float mu, float mu1,
                        float t, uint64_t size, float* v, const float* g_in, float* gv, float* w_io)
{
#pragma omp parallel for
    for (uint64_t i = 0; i < size; i++) {
        float g = g_in[i], w = w_io[i];
        switch (mode) {
        case ORC_SGD:       /* E/kernels.cu:2746-2757 */
            w_io[i] = w + alpha * (g - lambda * w - lambda1 * orc_sgn(w));
            break;
        case ORC_MOMENTUM: { /* :2798-2812 */
            float vv = mu * v[i] + alpha * (g - lambda * w - lambda1 * orc_sgn(w));
            v[i] = vv; w_io[i] = w + vv;
            break; }
        case ORC_ADAGRAD: {  /* :2854-2869 */
            g -= lambda * w + lambda1 * orc_sgn(w);
            float vv = v[i] + g * g;
            v[i] = vv; w_io[i] = w + alpha * g * (1.0f / sqrtf(fmaxf(0.000000001f, vv)));
            break; }
        case ORC_NESTEROV: { /* :3047-3062 */
            float vOld = v[i];
            float vNew = mu * vOld + alpha * (g - lambda * w - lambda1 * orc_sgn(w));
            v[i] = vNew; w_io[i] = w + vNew + mu * (vNew - vOld);
            break; }
        case ORC_RMSPROP: {  /* :3144-3159 */
            g -= lambda * w + lambda1 * orc_sgn(w);
            float vv = mu * v[i] + (1.0f - mu) * g * g;
            v[i] = vv; w_io[i] = w + alpha * g * (1.0f / sqrtf(fmaxf(0.000000001f, vv)));
            break; }
        case ORC_ADADELTA: { /* :2911-2930 */
            float vv = v[i], vg = gv[i];
            g -= lambda * w + lambda1 * orc_sgn(w);
            vg = mu * vg + (1.0f - mu) * g * g;
            float dw = sqrtf(fmaxf(0.000000001f, vv) / fmaxf(0.000000001f, vg)) * g;
            vv = mu * vv + (1.0f - mu) * dw * dw;
            v[i] = vv; gv[i] = vg; w_io[i] = w + dw;
            break; }
        case ORC_ADAM: {     /* :2976-2998; the kernel adds 1 to t */
            float dw = g, vdw = v[i], sdw = gv[i], tt = t + 1.0f;
            dw -= lambda * w + lambda1 * orc_sgn(w);
            vdw = mu * vdw + (1.0f - mu) * dw;
            sdw = mu1 * sdw + (1.0f - mu1) * dw * dw;
            v[i] = vdw; gv[i] = sdw;
            vdw /= 1.0f - powf(mu, tt);
            sdw /= 1.0f - powf(mu1, tt);
            dw = alpha * vdw / (sqrtf(sdw) + 1.0e-8f);
            w_io[i] = w + dw;
            break; }
        default: break;
        }
    }
}

void orc_update_biases(int mode, float alpha, float mu, float mu1, float t, uint32_t batch,
                       uint32_t width, const float* delta, float* v, float* gv, float* bias)
{
#pragma omp parallel for
    for (uint32_t c = 0; c < width; c++) {
        float sum = 0.0f;                          /* serial fp32 column sum, e.g. E/kernels.cu:2773-2781 */
        for (uint32_t b = 0; b < batch; b++) sum += delta[(size_t)b * width + c];
        sum /= (float)batch;
        switch (mode) {
        case ORC_SGD:      bias[c] = bias[c] - alpha * sum; break;                          /* :2766-2788 */
        case ORC_MOMENTUM: { float vv = mu * v[c] - alpha * sum; v[c] = vv; bias[c] += vv; break; }   /* :2821-2845 */
        case ORC_ADAGRAD:  { float vv = v[c] + sum * sum; v[c] = vv;
                             bias[c] -= alpha * sum * (1.0f / sqrtf(fmaxf(0.000000001f, vv))); break; } /* :2878-2902 */
        case ORC_NESTEROV: { float vOld = v[c]; float vNew = mu * vOld - alpha * sum; v[c] = vNew;
                             bias[c] += vNew + mu * (vNew - vOld); break; }                 /* :3071-3095 */
        case ORC_RMSPROP:  { float vv = mu * v[c] + (1.0f - mu) * sum * sum; v[c] = vv;
                             bias[c] -= alpha * sum * (1.0f / sqrtf(fmaxf(0.000000001f, vv))); break; } /* :3168-3192 */
        case ORC_ADADELTA: { float vv = v[c], vg = gv[c];                                   /* :2939-2967 */
                             vg = mu * vg + (1.0f - mu) * sum * sum;
                             float dw = sqrtf(fmaxf(0.000000001f, vv) / fmaxf(0.000000001f, vg)) * sum;
                             vv = mu * vv + (1.0f - mu) * dw * dw;
                             v[c] = vv; gv[c] = vg; bias[c] -= dw; break; }
        case ORC_ADAM:     { float vdw = v[c], sdw = gv[c], tt = t + 1.0f;                  /* :3007-3038 */
                             vdw = mu * vdw + (1.0f - mu) * sum;
                             sdw = mu1 * sdw + (1.0f - mu1) * sum * sum;
                             v[c] = vdw; gv[c] = sdw;
                             vdw /= 1.0f - powf(mu, tt);
                             sdw /= 1.0f - powf(mu1, tt);
                             bias[c] -= alpha * vdw / (sqrtf(sdw) + 1.0e-8f); break; }
        default: break;
        }
    }
}

double orc_regularization_error(float lambda, float lambda1, const float* w, uint64_t size)
{   /* E/kernels.cu:2719-2743: sum 0.5*lambda*w^2 + lambda1*|w| */
    double total = 0.0;
    float hl = 0.5f * lambda;
#pragma omp parallel for reduction(+ : total)
    for (uint64_t i = 0; i < size; i++) total += (double)(hl * w[i] * w[i] + lambda1 * fabsf(w[i]));
    return total;
}

/* ------------------------------------------------------------------ a13 */
typedef struct { float key; uint32_t idx; uint32_t val; } orc_kv;

static int orc_kv_cmp(const void* pa, const void* pb)
{
    const orc_kv* a = (const orc_kv*)pa; const orc_kv* b = (const orc_kv*)pb;
    if (a->key > b->key) return -1;
    if (a->key < b->key) return 1;
    return (a->idx > b->idx) - (a->idx < b->idx);
}

void orc_topk(const float* key, const uint32_t* inValue, uint32_t batch, uint32_t width, uint32_t k,
              const uint64_t* filterStart, const uint64_t* filterEnd, const uint32_t* filterIndex,
              float* outKey, uint32_t* outValue)
{
    /* Selection: E/kernels.cu:3472-3669 keeps the k largest keys in descending
     * order; candidates must be > the running k-th key and > -MAX_VALUE.
     * Filter: U/Filters.cpp:49-67 via U/NNRecsGenerator.cpp:138-147 multiplies
     * the listed scores by 0.  Canonical tie-break: ascending index. */
#pragma omp parallel
    {
        orc_kv* row = (orc_kv*)malloc((size_t)width * sizeof(orc_kv));
#pragma omp for schedule(dynamic, 1)
        for (uint32_t b = 0; b < batch; b++) {
            for (uint32_t i = 0; i < width; i++) {
                row[i].key = key[(size_t)b * width + i];
                row[i].idx = i;
                row[i].val = inValue ? inValue[(size_t)b * width + i] : i;
            }
            if (filterStart)
                for (uint64_t j = filterStart[b]; j < filterEnd[b]; j++)
                    if (filterIndex[j] < width) row[filterIndex[j]].key *= 0.0f;
            qsort(row, width, sizeof(orc_kv), orc_kv_cmp);
            for (uint32_t j = 0; j < k; j++) {
                if (j < width && row[j].key > -ORC_MAX_VALUE) {
                    outKey[(size_t)b * k + j] = row[j].key;
                    outValue[(size_t)b * k + j] = row[j].val;
                } else {
                    outKey[(size_t)b * k + j] = -ORC_MAX_VALUE;
                    outValue[(size_t)b * k + j] = 0;
                }
            }
        }
        free(row);
    }
}

/* ------------------------------------------------------------------ a15 */
void orc_shard_range(uint32_t N, uint32_t rank, uint32_t nranks, uint32_t* minX, uint32_t* maxX)
{   /* E/NNLayer.cpp:108-112: _minX = Nx*rank/P, _maxX = Nx*(rank+1)/P (size_t math) */
    *minX = (uint32_t)(((uint64_t)N * rank) / nranks);
    *maxX = (uint32_t)(((uint64_t)N * (rank + 1)) / nranks);
}

int orc_weight_outgoing_larger(uint32_t inputStride, uint32_t outputStride)
{   /* E/NNWeight.cpp:435-457 */
    return (uint64_t)outputStride * 3 > (uint64_t)inputStride * 2;
}

/* E/NNLayer.cpp:1685-1708, E/kernels.cu:4497-4537 */
void orc_dropout(int activation, float* unit, const float* random, uint32_t batch, uint32_t stride, float p, float eluAlpha, float seluLambda)
{
    const size_t size = (size_t)batch * stride;
    if (activation == ORC_ACT_ELU || activation == ORC_ACT_SELU) {
        const float lambda = (activation == ORC_ACT_SELU) ? seluLambda : 1.0f;
        const float alpha = -lambda * eluAlpha;
        const float q = 1.0f - p;
        const float a = 1.0f / sqrtf(q + alpha * alpha * p * q);
        const float b = -a * p * alpha;
        const float target = a * alpha + b;
        for (size_t i = 0; i < size; i++) unit[i] = (random[i] < p) ? target : a * unit[i] + b;
    } else {
        const float target = (activation == ORC_ACT_SIGMOID) ? 0.5f : 0.0f;
        const float scale = (target == 0.0f) ? 1.0f / (1.0f - p) : 1.0f;
        for (size_t i = 0; i < size; i++) unit[i] = (random[i] < p) ? target : scale * unit[i];
    }
}
