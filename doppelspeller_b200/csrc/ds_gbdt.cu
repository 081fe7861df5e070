// ds_gbdt.cu - f4: gradient-boosted tree inference over the [P, 66] feature matrix, on the device.
//
// Replaces `self.model.predict(xgb.DMatrix(features), ...)` of /root/reference/doppelspeller/predict.py:229-233
// (model trained by train.py:99-121: xgboost==0.90, requirements.txt:8; max_depth 5, <= 1000 rounds, objective
// reg:logistic).  xgboost is a third-party dependency that is neither vendored in the reference nor installed
// here, so its published prediction algorithm (xgboost 0.90 src/predictor/cpu_predictor.cc `PredValue`,
// include/xgboost/tree_model.h `GetNext`, src/objective/regression_loss.h `LogisticRegression::PredTransform`)
// is restated - parity with a real model file is unpinned:
//   per row:  psum = 0.0f; for every tree in boosting order: walk from the root - a NaN feature takes the node's
//             default child, otherwise `feature < split ? yes : no` (strict, float32) - and psum += leaf (float32);
//             margin = base_margin + psum;  reg:logistic / binary:logistic: 1 / (1 + exp(-margin)) in float32.
//
// One thread per row.  The trees of a model (<= 1000 x <= 63 nodes x 16 B = 1 MB) are streamed through shared
// memory in chunks shared by the 256 rows of a CTA; the CTA's 256 x 66 features are staged once, feature major
// with a row stride of 257 floats (conflict-free transposing stores; a lane reads bank (feature + lane) mod 32).
#include "ds_common.cuh"

namespace ds {

constexpr int GBDT_THREADS = 256;
constexpr int GBDT_CHUNK_NODES = 2560;   // 40 KB of nodes per chunk

static_assert(sizeof(ds_gbdt_node) == 16, "ds_gbdt_node must be 16 bytes");

template <bool STAGED>
__global__ void __launch_bounds__(GBDT_THREADS) k_gbdt_predict(const float *__restrict__ features, int64_t n_rows, int n_features,
                                                               const ds_gbdt_node *__restrict__ nodes,
                                                               const int32_t *__restrict__ tree_offsets, int n_trees, float base_margin,
                                                               int transform, float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char gbdt_smem[];
    uint4 *s_nodes = reinterpret_cast<uint4 *>(gbdt_smem);
    float *s_x = reinterpret_cast<float *>(gbdt_smem + (size_t)GBDT_CHUNK_NODES * 16);
    __shared__ int s_first_tree, s_last_tree;
    const int64_t row0 = (int64_t)blockIdx.x * GBDT_THREADS;
    const int64_t row = row0 + threadIdx.x;
    // x[f * x_stride]: this row's feature f
    const float *x = features + (row < n_rows ? row : 0) * (int64_t)n_features;
    int x_stride = 1;
    if (STAGED) {
        const int64_t rows_here = min((int64_t)GBDT_THREADS, n_rows - row0);
        const float *src = features + row0 * (int64_t)n_features;
        for (int64_t e = threadIdx.x; e < rows_here * n_features; e += GBDT_THREADS) {
            const int r = (int)(e / n_features), f = (int)(e % n_features);
            s_x[f * (GBDT_THREADS + 1) + r] = src[e];
        }
        x = s_x + threadIdx.x;
        x_stride = GBDT_THREADS + 1;
        __syncthreads();
    }
    float psum = 0.0f;
    int tree = 0;
    while (tree < n_trees) {
        // the next run of whole trees that fits the chunk (a single tree larger than the chunk is walked from global memory)
        if (threadIdx.x == 0) {
            int last = tree;
            const int first_node = tree_offsets[tree];
            while (last < n_trees && tree_offsets[last + 1] - first_node <= GBDT_CHUNK_NODES) ++last;
            s_first_tree = tree;
            s_last_tree = last;
        }
        __syncthreads();
        const int first = s_first_tree, last = s_last_tree;
        if (last == first) {
            const ds_gbdt_node *t = nodes + tree_offsets[first];
            int node = 0;
            ds_gbdt_node nd = t[0];
            while (nd.feature >= 0) {
                const float v = x[nd.feature * x_stride];
                node = isnan(v) ? nd.missing : (v < nd.value ? nd.yes : nd.no);
                nd = t[node];
            }
            psum = __fadd_rn(psum, nd.value);
            tree = first + 1;
            __syncthreads();
            continue;
        }
        const int node0 = tree_offsets[first], node1 = tree_offsets[last];
        const uint4 *src = reinterpret_cast<const uint4 *>(nodes + node0);
        for (int i = threadIdx.x; i < node1 - node0; i += GBDT_THREADS) s_nodes[i] = src[i];
        __syncthreads();
        if (row < n_rows) {
            // one step down a tree; a leaf stays where it is
            auto step = [&](uint4 raw, int root) -> uint4 {
                if ((int)raw.x < 0) return raw;
                const float v = x[raw.x * x_stride];
                const int yes = raw.z & 0xffff, no = raw.z >> 16, missing = raw.w & 0xffff;
                const int next = isnan(v) ? missing : (v < __uint_as_float(raw.y) ? yes : no);
                return s_nodes[root + next];
            };
            int t = first;
            // four trees are walked at once (independent dependency chains); their leaves are added in tree order
            for (; t + 4 <= last; t += 4) {
                const int root0 = tree_offsets[t] - node0, root1 = tree_offsets[t + 1] - node0;
                const int root2 = tree_offsets[t + 2] - node0, root3 = tree_offsets[t + 3] - node0;
                uint4 r0 = s_nodes[root0], r1 = s_nodes[root1], r2 = s_nodes[root2], r3 = s_nodes[root3];
                while ((int)(r0.x & r1.x & r2.x & r3.x) >= 0) {   // some walk has not reached its leaf (feature -1) yet
                    r0 = step(r0, root0);
                    r1 = step(r1, root1);
                    r2 = step(r2, root2);
                    r3 = step(r3, root3);
                }
                psum = __fadd_rn(psum, __uint_as_float(r0.y));
                psum = __fadd_rn(psum, __uint_as_float(r1.y));
                psum = __fadd_rn(psum, __uint_as_float(r2.y));
                psum = __fadd_rn(psum, __uint_as_float(r3.y));
            }
            for (; t < last; ++t) {
                const int root = tree_offsets[t] - node0;
                uint4 raw = s_nodes[root];
                while ((int)raw.x >= 0) raw = step(raw, root);
                psum = __fadd_rn(psum, __uint_as_float(raw.y));
            }
        }
        tree = last;
        __syncthreads();
    }
    if (row < n_rows) {
        float margin = __fadd_rn(base_margin, psum);
        if (transform == DS_GBDT_LOGISTIC) margin = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-margin)));
        out[row] = margin;
    }
}

}  // namespace ds

using namespace ds;

extern "C" {

int ds_gbdt_predict(const float *features, int64_t n_rows, int32_t n_features, const ds_gbdt_node *nodes,
                    const int32_t *tree_offsets, int32_t n_trees, float base_margin, int32_t transform, float *out,
                    void *stream_) {
    if (n_rows < 0 || n_features < 1 || n_trees < 0) return fail(DS_ERR_BAD_ARG, "bad n_rows / n_features / n_trees");
    if (n_rows > 0 && (features == nullptr || out == nullptr)) return fail(DS_ERR_BAD_ARG, "features / out is NULL");
    if (n_trees > 0 && (nodes == nullptr || tree_offsets == nullptr)) return fail(DS_ERR_BAD_ARG, "nodes / tree_offsets is NULL");
    if (transform != DS_GBDT_MARGIN && transform != DS_GBDT_LOGISTIC) return fail(DS_ERR_BAD_ARG, "unknown transform %d", transform);
    int n_devices = 0;
    if (cudaGetDeviceCount(&n_devices) != cudaSuccess || n_devices == 0) {
        cudaGetLastError();
        return fail(DS_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (n_rows == 0) return DS_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Workspace ws(stream);
    // validate the model on the host when it is host resident (child indexes inside their tree, features in range)
    std::vector<int32_t> h_offsets((size_t)n_trees + 1, 0);
    if (n_trees > 0) {
        if (is_device_pointer(tree_offsets)) {
            DS_CUDA(cudaMemcpyAsync(h_offsets.data(), tree_offsets, ((size_t)n_trees + 1) * 4, cudaMemcpyDeviceToHost, stream));
            DS_CUDA(cudaStreamSynchronize(stream));
        } else {
            memcpy(h_offsets.data(), tree_offsets, ((size_t)n_trees + 1) * 4);
        }
        if (h_offsets[0] != 0) return fail(DS_ERR_BAD_ARG, "tree_offsets must start at 0");
        for (int t = 0; t < n_trees; ++t) {
            const int64_t size = (int64_t)h_offsets[(size_t)t + 1] - h_offsets[(size_t)t];
            if (size < 1 || size > 65535) return fail(DS_ERR_BAD_ARG, "tree %d has %lld nodes (1..65535 allowed)", t, (long long)size);
        }
        if (!is_device_pointer(nodes)) {
            for (int t = 0; t < n_trees; ++t) {
                const int size = h_offsets[(size_t)t + 1] - h_offsets[(size_t)t];
                const ds_gbdt_node *tree = nodes + h_offsets[(size_t)t];
                for (int i = 0; i < size; ++i) {
                    if (tree[i].feature < 0) continue;
                    if (tree[i].feature >= n_features || tree[i].yes >= size || tree[i].no >= size || tree[i].missing >= size ||
                        tree[i].yes <= i || tree[i].no <= i || tree[i].missing <= i)
                        return fail(DS_ERR_BAD_ARG, "tree %d node %d: feature or child index out of range (children must follow their parent)", t, i);
                }
            }
        }
    }
    const int n_nodes = h_offsets[(size_t)n_trees];
    const float *d_features = nullptr;
    const ds_gbdt_node *d_nodes = nullptr;
    const int32_t *d_offsets = nullptr;
    float *d_out = nullptr;
    DS_CHECK(ws.stage_in(&d_features, features, (size_t)n_rows * n_features));
    if (n_nodes > 0) DS_CHECK(ws.stage_in(&d_nodes, nodes, (size_t)n_nodes));   // an empty model has no node to read (the kernel never looks)
    DS_CHECK(ws.stage_in(&d_offsets, tree_offsets, (size_t)n_trees + 1));
    DS_CHECK(ws.stage_out(&d_out, out, (size_t)n_rows));
    const int32_t zero = 0;
    if (n_trees == 0) {
        int32_t *d_zero = nullptr;
        DS_CHECK(ws.alloc(&d_zero, 1));
        DS_CUDA(cudaMemcpyAsync(d_zero, &zero, 4, cudaMemcpyHostToDevice, stream));
        d_offsets = d_zero;
    }
    const size_t node_bytes = (size_t)GBDT_CHUNK_NODES * 16;
    const size_t staged_bytes = node_bytes + (size_t)n_features * (GBDT_THREADS + 1) * 4;
    const unsigned blocks = (unsigned)ceil_div(n_rows, GBDT_THREADS);
    if (staged_bytes <= 112 * 1024) {   // two CTAs per SM
        DS_CHECK(ensure_dynamic_smem(reinterpret_cast<const void *>(&k_gbdt_predict<true>), staged_bytes));
        k_gbdt_predict<true><<<blocks, GBDT_THREADS, staged_bytes, stream>>>(d_features, n_rows, n_features, d_nodes, d_offsets, n_trees,
                                                                            base_margin, transform, d_out);
    } else {
        k_gbdt_predict<false><<<blocks, GBDT_THREADS, node_bytes, stream>>>(d_features, n_rows, n_features, d_nodes, d_offsets, n_trees,
                                                                           base_margin, transform, d_out);
    }
    DS_LAUNCHED("k_gbdt_predict");
    return ws.finish_outputs();
}

}  // extern "C"
