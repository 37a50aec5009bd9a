// shim_smoke.cpp -- one translation unit written the way the REFERENCE's callers are written (NNLayer / NNWeight / NNDataSet
// bodies: free functions of E/kernels.h on raw device pointers), compiled against include/dsstne_b200_kernels.hpp instead of
// E/kernels.h and RUN: a miniature training step of a sparse autoencoder (16 examples, 64 -> 8 -> 64) through the shim names,
// checked against a host loop.  Exit code 0 and "shim ok" = every shim launcher it touches forwards correctly.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/dsstne_b200_kernels.hpp"

template <typename T>
static T* upload(const std::vector<T>& v)
{
    T* d = nullptr;
    cudaMalloc(&d, (v.empty() ? 1 : v.size()) * sizeof(T));
    if (!v.empty()) cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}
template <typename T>
static std::vector<T> download(const T* d, size_t n)
{
    std::vector<T> v(n);
    cudaMemcpy(v.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost);
    return v;
}
static void require(bool ok, const char* what)
{
    if (!ok) { fprintf(stderr, "shim_smoke: FAILED %s\n", what); exit(1); }
}

int main()
{
    dsb200_ctx* ctx = nullptr;
    if (dsb200_ctx_create(&ctx, 0)) { fprintf(stderr, "shim_smoke: no sm_100 GPU (there is no CPU fallback)\n"); return 2; }
    dsb200k::bind(ctx);
    const uint32_t B = 16, N = 64, S = 8, K = 4;
    // CSR: example b holds columns b, b + 7, b + 20 (mod N)
    std::vector<uint64_t> start(B), end(B);
    std::vector<uint32_t> index;
    for (uint32_t b = 0; b < B; b++) {
        start[b] = index.size();
        uint32_t c[3] = {b % N, (b + 7) % N, (b + 20) % N};
        for (int i = 0; i < 3; i++) index.push_back(c[i]);
        end[b] = index.size();
    }
    std::vector<float> W1(N * S), b1(S), W2(S * N), b2(N);
    for (size_t i = 0; i < W1.size(); i++) W1[i] = 0.01f * (float)((int)(i * 37 % 23) - 11);
    for (size_t i = 0; i < W2.size(); i++) W2[i] = 0.02f * (float)((int)(i * 17 % 19) - 9);
    for (uint32_t i = 0; i < S; i++) b1[i] = 0.1f * (float)i - 0.3f;
    for (uint32_t i = 0; i < N; i++) b2[i] = -0.5f;
    uint64_t* dStart = upload(start); uint64_t* dEnd = upload(end); uint32_t* dIndex = upload(index);
    float* dW1 = upload(W1); float* db1 = upload(b1);
    float *dH, *dO, *dDeltaO, *dDeltaH, *dG1, *dG2;
    cudaMalloc(&dH, B * S * 4); cudaMalloc(&dO, B * N * 4); cudaMalloc(&dDeltaO, B * N * 4); cudaMalloc(&dDeltaH, B * S * 4);
    cudaMalloc(&dG1, N * S * 4); cudaMalloc(&dG2, S * N * 4);

    // ---- forward, as NNLayer::ForwardPropagateFullyConnected writes it (E/NNLayer.cpp:1002-1157)
    kClearUnit(dH, db1, S, B);
    kCalculateSparseZ(0, B, S, dW1, dStart, dEnd, dIndex, nullptr, dH, 1.0f);
    kCalculateSigmoidActivation(dH, (uint64_t)B * S);
    std::vector<float> H = download(dH, B * S);
    for (uint32_t b = 0; b < B; b++)
        for (uint32_t s = 0; s < S; s++) {
            float z = b1[s];
            for (uint64_t j = start[b]; j < end[b]; j++) z += W1[index[j] * S + s];
            require(std::fabs(H[b * S + s] - 1.0f / (1.0f + std::exp(-z))) < 1e-5f, "kClearUnit + kCalculateSparseZ + kCalculateSigmoidActivation");
        }
    // output layer on the host-checked hidden units: plain loops stand in for cublasSgemm (the GEMM entry points are C-ABI only)
    std::vector<float> O(B * N);
    for (uint32_t b = 0; b < B; b++)
        for (uint32_t n = 0; n < N; n++) {
            float z = b2[n];
            for (uint32_t s = 0; s < S; s++) z += H[b * S + s] * W2[s * N + n];
            O[b * N + n] = z;
        }
    cudaMemcpy(dO, O.data(), O.size() * 4, cudaMemcpyHostToDevice);
    kCalculateSigmoidActivation(dO, (uint64_t)B * N);
    // ---- loss and delta over the sparse targets (E/NNLayer.cpp:1710-1806)
    const float loss = kCalculateSparseScaledMarginalCrossEntropyError(0, B, N, dO, dStart, dEnd, dIndex, nullptr, false);
    kCalculateSparseScaledMarginalCrossEntropyOutputDelta(0 /* Sigmoid */, 0, B, N, dO, dDeltaO, dStart, dEnd, dIndex, nullptr, false);
    std::vector<float> A = download(dO, B * N), D = download(dDeltaO, B * N);
    dsb200_params P; dsb200_params_default(&P);
    double want = 0.0;
    for (uint32_t b = 0; b < B; b++)
        for (uint32_t n = 0; n < N; n++) {
            bool nz = false;
            for (uint64_t j = start[b]; j < end[b]; j++) nz = nz || index[j] == n;
            const float a = A[b * N + n];
            float d;
            if (nz) { d = (a < P.SMCE_oneTarget) ? P.SMCE_oneScale * (a - 1.0f) : 0.0f; if (a < P.SMCE_oneTarget) want += -P.SMCE_oneScale * std::log(std::fmax(1e-12f, a)); }
            else    { d = (a > P.SMCE_zeroTarget) ? P.SMCE_zeroScale * a : 0.0f; if (a > P.SMCE_zeroTarget) want += -P.SMCE_zeroScale * std::log(std::fmax(1e-12f, 1.0f - a)); }
            require(std::fabs(D[b * N + n] - d) < 1e-5f, "kCalculateSparseScaledMarginalCrossEntropyOutputDelta");
        }
    require(std::fabs(loss - (float)want) < 1e-4f * (float)std::fabs(want), "kCalculateSparseScaledMarginalCrossEntropyError");
    // ---- backward through the input weight: transposed matrix + sparse gradient (E/NNLayer.cpp:920-957, 2219)
    std::vector<float> dH0(B * S);
    for (size_t i = 0; i < dH0.size(); i++) dH0[i] = 0.001f * (float)((int)(i % 13) - 6);
    cudaMemcpy(dDeltaH, dH0.data(), dH0.size() * 4, cudaMemcpyHostToDevice);
    kCalculateHadamardProduct(0 /* Sigmoid */, (uint64_t)B * S, 1.0f, dH, dDeltaH, 0.0f, 0.0f, 0.0f);
    std::vector<uint32_t> tStart(N), cnt(N, 0);
    for (uint32_t c : index) cnt[c]++;
    uint32_t run = 0;
    for (uint32_t c = 0; c < N; c++) { tStart[c] = run; run += (cnt[c] + 31) / 32 * 32; }
    uint32_t* dTStart = upload(tStart); uint32_t* dTEnd = upload(tStart);       // End <- Start, as the reference's caller does (E/NNTypes.h:576)
    uint32_t* dTIndex = nullptr; cudaMalloc(&dTIndex, (run ? run : 1) * 4);
    kCalculateSparseTransposedMatrix(0, B, dStart, dEnd, dIndex, nullptr, dTEnd, dTIndex, nullptr, N);
    kCalculateSparseTransposedWeightGradient(-1.0f / B, 0.0f, N, S, dTStart, dTEnd, dTIndex, dDeltaH, dG1);
    std::vector<float> G1 = download(dG1, N * S), dHh = download(dDeltaH, B * S);
    for (uint32_t c = 0; c < N; c++)
        for (uint32_t s = 0; s < S; s++) {
            double g = 0.0;
            for (uint32_t b = 0; b < B; b++)
                for (uint64_t j = start[b]; j < end[b]; j++) if (index[j] == c) g += dHh[b * S + s];
            require(std::fabs(G1[c * S + s] - (float)(-g / B)) < 1e-6f, "kCalculateSparseTransposedMatrix + kCalculateSparseTransposedWeightGradient");
        }
    // ---- optimizer (E/NNWeight.cpp:729-794)
    kSGDUpdateWeights(0.5f, 0.0f, 0.0f, (uint64_t)N * S, dG1, dW1);
    kSGDUpdateBiases(0.5f, B, S, dDeltaH, db1);
    std::vector<float> W1n = download(dW1, N * S), b1n = download(db1, S);
    for (size_t i = 0; i < W1.size(); i++) require(std::fabs(W1n[i] - (W1[i] + 0.5f * G1[i])) < 1e-6f, "kSGDUpdateWeights");
    for (uint32_t s = 0; s < S; s++) {
        double m = 0.0;
        for (uint32_t b = 0; b < B; b++) m += dHh[b * S + s];
        require(std::fabs(b1n[s] - (b1[s] - 0.5f * (float)(m / B))) < 1e-6f, "kSGDUpdateBiases");
    }
    // ---- prediction: top-K of the output scores (E/NNNetwork.cpp:1792-1822), 3-arg and float-valued 4-arg forms
    float* dKey; uint32_t* dVal; float* dValF; float* dKey2;
    cudaMalloc(&dKey, B * K * 4); cudaMalloc(&dVal, B * K * 4); cudaMalloc(&dValF, B * K * 4); cudaMalloc(&dKey2, B * K * 4);
    kCalculateTopK(dO, dKey, dVal, B, N, K);
    std::vector<float> key = download(dKey, B * K);
    std::vector<uint32_t> val = download(dVal, B * K);
    for (uint32_t b = 0; b < B; b++) {
        float best = -1.0f; uint32_t arg = 0;
        for (uint32_t n = 0; n < N; n++) if (A[b * N + n] > best) { best = A[b * N + n]; arg = n; }
        require(key[b * K] == best && val[b * K] == arg, "kCalculateTopK (3-arg)");
        for (uint32_t k = 1; k < K; k++) require(key[b * K + k] <= key[b * K + k - 1], "kCalculateTopK order");
    }
    kCalculateTopK(dO, dDeltaO, dKey2, dValF, B, N, K);                          // E/kernels.h:42: (scores, values riding along) in, (keys, values) out
    std::vector<float> key2 = download(dKey2, B * K), valF = download(dValF, B * K);
    for (uint32_t b = 0; b < B; b++) {
        if (!(key2[b * K] == key[b * K] && valF[b * K] == D[b * N + val[b * K]]))
            fprintf(stderr, "row %u: key %.9g vs %.9g, value %.9g vs %.9g (column %u)\n", b, key2[b * K], key[b * K], valF[b * K], D[b * N + val[b * K]], val[b * K]);
        require(key2[b * K] == key[b * K] && valF[b * K] == D[b * N + val[b * K]], "kCalculateTopK (4-arg, float values)");
    }
    require(dsb200_ctx_sync(ctx) == 0, "dsb200_ctx_sync");
    dsb200_ctx_destroy(ctx);
    printf("shim ok: %d launches through the E/kernels.h names\n", (int)dsb200_launch_count());
    return 0;
}
