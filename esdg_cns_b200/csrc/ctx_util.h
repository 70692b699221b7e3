// Small host helpers shared by the translation units of libesdg_b200.so.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "esdg_ctx.h"

namespace esdg_host {

inline int fail(esdg_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? ESDG_ENOMEM : ESDG_ECUDA,             \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                         \
    } while (0)

template <class T>
inline int dev_alloc(esdg_ctx* ctx, T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return fail(ctx, ESDG_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    ctx->allocs.push_back(q);
    *p = (T*)q;
    return ESDG_OK;
}

template <class T>
inline int dev_upload(esdg_ctx* ctx, T** p, const std::vector<T>& h) {
    int rc = dev_alloc(ctx, p, h.size());
    if (rc) return rc;
    if (!h.empty()) CK(cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return ESDG_OK;
}

// dense column-major -> ELL with the reference's drop tolerance
inline int build_ell(esdg_ctx* ctx, const HostMat& A, double tol, DevEll* out) {
    const int rows = (int)A.rows, cols = (int)A.cols;
    std::vector<int> cnt(rows, 0);
    int width = 0;
    for (int i = 0; i < rows; ++i) {
        for (int j = 0; j < cols; ++j)
            if (std::fabs(A.at(i, j)) > tol) cnt[i]++;
        width = std::max(width, cnt[i]);
    }
    width = std::max(width, 1);
    std::vector<int> col((size_t)rows * width, 0);
    std::vector<double> val((size_t)rows * width, 0.0);
    for (int i = 0; i < rows; ++i) {
        int k = 0;
        for (int j = 0; j < cols; ++j)
            if (std::fabs(A.at(i, j)) > tol) {
                col[(size_t)i * width + k] = j;
                val[(size_t)i * width + k] = A.at(i, j);
                ++k;
            }
    }
    out->rows = rows;
    out->width = width;
    int rc;
    if ((rc = dev_upload(ctx, &out->cnt, cnt))) return rc;
    if ((rc = dev_upload(ctx, &out->col, col))) return rc;
    if ((rc = dev_upload(ctx, &out->val, val))) return rc;
    return ESDG_OK;
}

// dense column-major -> row-major, zero-padded to (ceil8(rows), ceil4(cols)), times `scale`; rows [r0, r0+nr)
inline int build_dense(esdg_ctx* ctx, const HostMat& A, double scale, int r0, int nr, esdg::Dense* out) {
    const int K = (int)A.cols, Mp = (nr + 7) / 8 * 8, Kp = (K + 3) / 4 * 4;
    std::vector<double> h((size_t)Mp * Kp, 0.0);
    for (int i = 0; i < nr; ++i)
        for (int j = 0; j < K; ++j) h[(size_t)i * Kp + j] = scale * A.at(r0 + i, j);
    double* d = nullptr;
    int rc = dev_upload(ctx, &d, h);
    if (rc) return rc;
    *out = esdg::Dense{d, nr, K, Mp, Kp};
    return ESDG_OK;
}

inline const HostMat* find(const std::map<std::string, HostMat>& m, const char* name) {
    auto it = m.find(name);
    return it == m.end() ? nullptr : &it->second;
}

inline void tick(esdg_ctx* ctx, const char* name) {
    ctx->launches++;
    if (ctx->timing) {
        KernelTimer t;
        t.name = name;
        cudaEventCreate(&t.a);
        cudaEventCreate(&t.b);
        cudaEventRecord(t.a, ctx->stream);
        ctx->timers.push_back(t);
    }
}
inline void tock(esdg_ctx* ctx) {
    if (ctx->timing) cudaEventRecord(ctx->timers.back().b, ctx->stream);
}


}  // namespace esdg_host
