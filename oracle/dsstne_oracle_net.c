/*
 * dsstne_oracle_net.c -- whole-network CPU restatement for sparse-input /
 * sparse-target fully-connected networks (single process), built from the
 * per-kernel oracle functions.  TEST INFRASTRUCTURE ONLY; see dsstne_oracle.h.
 *
 * Follows, for one minibatch, the body of NNNetwork::Train's loop
 * (E/NNNetwork.cpp:1601-1650):
 *   PredictTrainingBatch  E/NNNetwork.cpp:1431  -> NNLayer::LoadTrainingBatch E/NNLayer.cpp:910
 *                                              -> ForwardPropagateFullyConnected E/NNLayer.cpp:994-1160
 *   CalculateError        E/NNNetwork.cpp:1708  -> NNLayer::CalculateError E/NNLayer.cpp:1710
 *   BackPropagate         E/NNNetwork.cpp:1747  -> CalculateOutputDelta E/NNLayer.cpp:1762,
 *                                                  BackPropagateFullyConnected E/NNLayer.cpp:2121-2316
 *   _batches++ ; UpdateWeights E/NNNetwork.cpp:1770 -> NNWeight::UpdateWeights E/NNWeight.cpp:718
 */
#include "dsstne_oracle.h"

#include <stdlib.h>
#include <string.h>

static float* orc_falloc(size_t n) { return (float*)calloc(n ? n : 1, sizeof(float)); }

orc_network* orc_net_create(int nWeights, const uint32_t* sizes, const int* activations,
                            int errorFunction, int trainingMode, uint32_t maxBatch)
{
    if (nWeights < 1 || nWeights > ORC_MAX_WEIGHTS) return NULL;
    orc_network* net = (orc_network*)calloc(1, sizeof(orc_network));
    net->nWeights = nWeights;
    net->errorFunction = errorFunction;
    net->trainingMode = trainingMode;
    net->maxBatch = maxBatch;
    orc_params_default(&net->params);
    for (int l = 0; l <= nWeights; l++) {
        net->size[l] = sizes[l];
        net->activation[l] = activations ? activations[l] : ORC_ACT_SIGMOID;
    }
    for (int i = 0; i < nWeights; i++) {
        size_t n = (size_t)sizes[i] * sizes[i + 1];
        net->W[i] = orc_falloc(n);  net->dW[i] = orc_falloc(n);
        net->vW[i] = orc_falloc(n); net->gvW[i] = orc_falloc(n);
        net->b[i] = orc_falloc(sizes[i + 1]);
        net->vb[i] = orc_falloc(sizes[i + 1]); net->gvb[i] = orc_falloc(sizes[i + 1]);
    }
    for (int l = 1; l <= nWeights; l++) {
        net->unit[l] = orc_falloc((size_t)maxBatch * sizes[l]);
        net->delta[l] = orc_falloc((size_t)maxBatch * sizes[l]);
    }
    return net;
}

void orc_net_destroy(orc_network* net)
{
    if (!net) return;
    for (int i = 0; i < net->nWeights; i++) {
        free(net->W[i]); free(net->dW[i]); free(net->vW[i]); free(net->gvW[i]);
        free(net->b[i]); free(net->vb[i]); free(net->gvb[i]);
    }
    for (int l = 1; l <= net->nWeights; l++) { free(net->unit[l]); free(net->delta[l]); }
    free(net->tStart); free(net->tEnd); free(net->tIndex); free(net->tData);
    free(net);
}

void orc_net_set_input(orc_network* net, const orc_csr* in, uint32_t examples, uint32_t uniqueExamples, uint32_t batch)
{
    /* NNDataSet::GenerateSparseTransposedMatrix, E/NNTypes.cpp:1520-1570 */
    uint32_t N = net->size[0];
    free(net->tStart); free(net->tEnd); free(net->tIndex); free(net->tData);
    net->tStart = (uint32_t*)calloc(N, sizeof(uint32_t));
    net->tEnd = (uint32_t*)calloc(N, sizeof(uint32_t));
    net->tCapacity = orc_transposed_capacity(in, examples, uniqueExamples, N, batch, net->tStart);
    net->tIndex = (uint32_t*)calloc(net->tCapacity ? net->tCapacity : 1, sizeof(uint32_t));
    net->tData = (in->sparseData || in->dataWeight) ? orc_falloc(net->tCapacity) : NULL;
}

void orc_net_forward(orc_network* net, const orc_csr* in, uint32_t position, uint32_t batch, int training)
{
    for (int l = 1; l <= net->nWeights; l++) {
        uint32_t S = net->size[l];
        orc_clear_unit(net->unit[l], net->b[l - 1], S, batch);                       /* E/NNLayer.cpp:1009 */
        if (l == 1)                                                                   /* E/NNLayer.cpp:1046-1055 */
            orc_sparse_z(&net->params, in, position, batch, S, net->W[0], net->unit[1], 1.0f,
                         training && net->denoising);
        else                                                                          /* E/NNLayer.cpp:1057-1086 */
            orc_gemm_fwd(batch, net->size[l - 1], S, net->unit[l - 1], net->W[l - 1], 1.0f, net->unit[l]);
        orc_activation(net->activation[l], net->unit[l], batch, S, 0.0f, 0.0f, 0.0f);  /* E/NNLayer.cpp:1157 */
        if (training && l < net->nWeights && net->pDropout[l] > 0.0f && net->dropoutRandom[l])    /* E/NNLayer.cpp:1160 */
            orc_dropout(net->activation[l], net->unit[l], net->dropoutRandom[l], batch, S, net->pDropout[l], 1.0f, 1.050701f);
    }
}

static void orc_net_load_training_batch(orc_network* net, const orc_csr* in, uint32_t position, uint32_t batch)
{
    /* NNLayer::LoadTrainingBatch (fast sparse), E/NNLayer.cpp:916-925:
     * End <- Start (E/NNTypes.h:576) then the transposed build. */
    memcpy(net->tEnd, net->tStart, (size_t)net->size[0] * sizeof(uint32_t));
    orc_sparse_transpose(&net->params, in, position, batch, net->denoising, net->tEnd, net->tIndex, net->tData);
}

double orc_net_loss(orc_network* net, const orc_csr* in, const orc_csr* out, uint32_t position, uint32_t batch)
{
    int L = net->nWeights;
    orc_net_forward(net, in, position, batch, 1);
    return orc_sparse_loss(&net->params, out, net->errorFunction, net->activation[L], position, batch,
                           net->size[L], net->unit[L], 0);
}

static void orc_net_backprop(orc_network* net, const orc_csr* out, uint32_t position, uint32_t batch)
{
    int L = net->nWeights;
    orc_sparse_output_delta(&net->params, out, net->errorFunction, net->activation[L], position, batch,
                            net->size[L], net->unit[L], net->delta[L], 0, 0.0f, 0.0f, 0.0f);
    for (int l = L; l >= 1; l--) {
        uint32_t S = net->size[l];
        if (l < L) {                                                     /* Hidden: E/NNLayer.cpp:2127-2138 */
            if (net->sparsePenalty[l] && net->sparsenessPenalty_beta > 0.0f)
                orc_sparseness_penalty(batch, S, net->unit[l], net->delta[l],
                                       net->sparsenessPenalty_p, net->sparsenessPenalty_beta);
            orc_hadamard(net->activation[l], (uint64_t)batch * S, 1.0f / (1.0f - net->pDropout[l]), net->unit[l], net->delta[l],
                         0.0f, 0.0f, 0.0f);                          /* scale: E/NNLayer.cpp:2137 */
        }
        float galpha = -1.0f / (float)batch;                             /* E/NNLayer.cpp:2213 (sharingCount 1) */
        if (l == 1)                                                      /* E/NNLayer.cpp:2217-2220 */
            orc_sparse_wgrad(&net->params, galpha, 0.0f, net->size[0], S, net->tStart, net->tEnd,
                             net->tIndex, net->tData, net->delta[1], net->dW[0]);
        else {
            orc_gemm_dw(batch, net->size[l - 1], S, galpha, net->unit[l - 1], net->delta[l], 0.0f, net->dW[l - 1]);
            orc_gemm_dx(batch, net->size[l - 1], S, net->delta[l], net->W[l - 1], 0.0f, net->delta[l - 1]);
        }
    }
}

void orc_net_backward(orc_network* net, const orc_csr* in, const orc_csr* out, uint32_t position, uint32_t batch)
{
    orc_net_load_training_batch(net, in, position, batch);
    orc_net_forward(net, in, position, batch, 1);
    orc_net_backprop(net, out, position, batch);
}

double orc_net_train_step(orc_network* net, const orc_csr* in, const orc_csr* out, uint32_t position,
                          uint32_t batch, float alpha, float lambda, float lambda1, float mu, float mu1,
                          double* reg)
{
    int L = net->nWeights;
    orc_net_load_training_batch(net, in, position, batch);
    orc_net_forward(net, in, position, batch, 1);
    double err = orc_sparse_loss(&net->params, out, net->errorFunction, net->activation[L], position, batch,
                                 net->size[L], net->unit[L], 0);
    double r = 0.0;
    if (lambda != 0.0f || lambda1 != 0.0f)                               /* E/NNNetwork.cpp:1724-1730 */
        for (int i = 0; i < L; i++)
            r += orc_regularization_error(lambda, lambda1, net->W[i], (uint64_t)net->size[i] * net->size[i + 1]);
    if (reg) *reg = r;
    orc_net_backprop(net, out, position, batch);
    net->batches++;                                                      /* E/NNNetwork.cpp:1647 */
    for (int i = L - 1; i >= 0; i--) {                                   /* E/NNNetwork.cpp:1777-1780 */
        orc_update_weights(net->trainingMode, alpha, lambda, lambda1, mu, mu1, (float)net->batches,
                           (uint64_t)net->size[i] * net->size[i + 1], net->vW[i], net->dW[i], net->gvW[i], net->W[i]);
        orc_update_biases(net->trainingMode, alpha, mu, mu1, (float)net->batches, batch, net->size[i + 1],
                          net->delta[i + 1], net->vb[i], net->gvb[i], net->b[i]);
    }
    return err;
}
