// rf_induction.cu -- random-forest tree induction statistics on the device.
//
// The reference grows a tree node by node (random_forest.rs:362-408): per node FeatureStats
// (normalizers.rs:13-36: min / max per feature), then per feature k-1 evenly spaced thresholds
// between min and max, each scored from the labels on its two sides (:211-286).  Here a tree
// grows LEVEL by level; the instances sampled for the tree carry the id of the node they sit in,
// and two passes over (instance, feature) pairs produce, for every active node at once,
//   pass 1   min / max of every sampled feature, and the node's label statistics
//   pass 2   per (node, feature, bucket between consecutive thresholds): count, positives,
//            sum and sum of squares of the labels
// from which the host scores every candidate split (prefix sums over buckets) exactly as the
// reference defines them, picks the winners and sends back a partition table.  Label sums are
// integers (labels in units of 2^-12), so the statistics -- and the forest -- do not depend on
// the order in which threads add.  X is read feature-major from the resident matrix: threads of
// a warp read consecutive positions of one feature row.
#include "device_common.cuh"

namespace {

constexpr int kChunk = 2048;   // sampled instances per CTA
constexpr int kRfThreads = 256;
constexpr double kGainScale = 4096.0;  // labels are accumulated in units of 2^-12 (FR_RF_GAIN_BITS)

__device__ __forceinline__ int ford(float v) {  // order-preserving float -> int
    const int i = __float_as_int(v);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float fdro(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct RfLevel {
    const float *x;
    size_t ld;
    const float *gain;        // by position
    const uint32_t *len_pos;  // nullptr, or by position: feature ids >= len_pos[p] are missing in that row
    const uint32_t *present_bits;  // nullptr, or by position: present_words words, bit f = the row carries feature f
    uint32_t present_words;
    unsigned *f_present;      // [n_active][F] instances that carry the feature (written when rows can miss features)
    const uint32_t *samp_pos; // [m] position of every sampled instance
    const int *node_of;       // [m] active node of the instance, -1 = settled in a leaf
    uint32_t m;
    const uint32_t *feats;    // [F]
    uint32_t F, n_active, k;
    int *fmin, *fmax;                 // [n_active][F] ordered-int min / max
    unsigned long long *node_n;       // [n_active]
    long long *node_sum;              // [n_active] labels, 2^-12 units
    int *gmin, *gmax;                 // [n_active] ordered-int label min / max
    unsigned *b_n, *b_pos;            // [n_active][F][k]
    long long *b_sum, *b_sq;          // [n_active][F][k]  2^-12 and 2^-24 units
};

// pass 1: grid (chunks, F)
__global__ void __launch_bounds__(kRfThreads) rf_minmax_kernel(RfLevel L) {
    extern __shared__ int sm[];
    int *smn = sm, *smx = sm + L.n_active;
    const uint32_t f = blockIdx.y;
    const bool labels = f == 0;
    int *sgmn = smx + L.n_active, *sgmx = sgmn + L.n_active;
    unsigned *scnt = (unsigned *)(sgmx + L.n_active);
    long long *ssum = (long long *)(scnt + L.n_active + (L.n_active & 1));
    unsigned *sfp = (unsigned *)(ssum + L.n_active);  // instances of the node that carry this feature
    const bool sparse = L.len_pos != nullptr || L.present_bits != nullptr;
    const uint32_t fid = L.feats[f];
    for (uint32_t a = threadIdx.x; a < L.n_active; a += blockDim.x) {
        smn[a] = INT_MAX;
        smx[a] = INT_MIN;
        sfp[a] = 0u;
        if (labels) {
            sgmn[a] = INT_MAX;
            sgmx[a] = INT_MIN;
            scnt[a] = 0u;
            ssum[a] = 0ll;
        }
    }
    __syncthreads();
    const float *__restrict__ row = L.x + (size_t)L.feats[f] * L.ld;
    const uint32_t i0 = blockIdx.x * kChunk, i1 = min(L.m, i0 + kChunk);
    // Near the root whole warps sit in one node: such a warp reduces in registers and issues one
    // shared-memory atomic per quantity instead of 32 colliding ones.
    for (uint32_t base = i0; base < i1; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const int nid = i < i1 ? L.node_of[i] : -1;
        int uniform = 0;
        __match_all_sync(0xffffffffu, nid, &uniform);
        if (uniform && nid < 0) continue;
        const uint32_t p = nid >= 0 ? L.samp_pos[i] : 0u;
        // normalizers.rs:21-27: a row that does not carry the feature is left out of its min / max
        bool present = nid >= 0;
        if (present && L.present_bits)
            present = (__ldg(L.present_bits + (size_t)p * L.present_words + (fid >> 5)) >> (fid & 31u)) & 1u;
        else if (present && L.len_pos)
            present = fid < __ldg(L.len_pos + p);
        const int o = nid >= 0 ? ford(__ldg(row + p)) : 0;
        const float g = (labels && nid >= 0) ? __ldg(L.gain + p) : 0.f;
        const int og = ford(g);
        const int y = (int)__double2ll_rn((double)g * kGainScale);  // |label| <= 16: fits easily
        if (uniform) {
            const int wmn = __reduce_min_sync(0xffffffffu, present ? o : INT_MAX);
            const int wmx = __reduce_max_sync(0xffffffffu, present ? o : INT_MIN);
            const unsigned npresent = (unsigned)__popc(__ballot_sync(0xffffffffu, present));
            int gmn = 0, gmx = 0, ys = 0;
            if (labels) {
                gmn = __reduce_min_sync(0xffffffffu, og);
                gmx = __reduce_max_sync(0xffffffffu, og);
                ys = __reduce_add_sync(0xffffffffu, y);
            }
            if ((threadIdx.x & 31) == 0) {
                if (npresent) {
                    atomicMin(&smn[nid], wmn);
                    atomicMax(&smx[nid], wmx);
                    if (sparse) atomicAdd(&sfp[nid], npresent);
                }
                if (labels) {
                    atomicMin(&sgmn[nid], gmn);
                    atomicMax(&sgmx[nid], gmx);
                    atomicAdd(&scnt[nid], 32u);
                    atomicAdd((unsigned long long *)&ssum[nid], (unsigned long long)(long long)ys);
                }
            }
        } else if (nid >= 0) {
            if (present) {
                atomicMin(&smn[nid], o);
                atomicMax(&smx[nid], o);
                if (sparse) atomicAdd(&sfp[nid], 1u);
            }
            if (labels) {
                atomicMin(&sgmn[nid], og);
                atomicMax(&sgmx[nid], og);
                atomicAdd(&scnt[nid], 1u);
                atomicAdd((unsigned long long *)&ssum[nid], (unsigned long long)(long long)y);
            }
        }
    }
    __syncthreads();
    for (uint32_t a = threadIdx.x; a < L.n_active; a += blockDim.x) {
        if (smn[a] != INT_MAX) {
            atomicMin(&L.fmin[(size_t)a * L.F + f], smn[a]);
            atomicMax(&L.fmax[(size_t)a * L.F + f], smx[a]);
        }
        if (sparse && sfp[a]) atomicAdd(&L.f_present[(size_t)a * L.F + f], sfp[a]);
        if (labels && scnt[a]) {
            atomicMin(&L.gmin[a], sgmn[a]);
            atomicMax(&L.gmax[a], sgmx[a]);
            atomicAdd(&L.node_n[a], (unsigned long long)scnt[a]);
            atomicAdd((unsigned long long *)&L.node_sum[a], (unsigned long long)ssum[a]);
        }
    }
}

// pass 2: grid (chunks, F).  Bucket b of a value = number of thresholds it is not below, i.e. the
// first i (1-based) with v < i/k * range + min is bucket i-1; values above every threshold fall in
// bucket k-1.  The threshold expression is random_forest.rs:231-234, evaluated in f64 with a
// separate multiply and add, as on the host.
template <bool SMEM>
__global__ void __launch_bounds__(kRfThreads) rf_bucket_kernel(RfLevel L) {
    extern __shared__ long long sm64[];
    const uint32_t f = blockIdx.y, k = L.k;
    const size_t cells = (size_t)L.n_active * k;
    // Shared-memory cell = four 32-bit words (64-bit shared-memory atomics are several times
    // slower than 32-bit ones here): a CTA sees at most kChunk = 2048 instances and |y| <= 2^16, so
    //   [0] instances | instances with label > 0 << 16    [1] sum of y (signed)
    //   [2] sum of (y^2 & 0xffff)                          [3] sum of (y^2 >> 16)
    // all stay below 2^31.
    unsigned *scell = (unsigned *)sm64;
    if (SMEM) {
        for (size_t c = threadIdx.x; c < 4 * cells; c += blockDim.x) scell[c] = 0u;
        __syncthreads();
    }
    const float *__restrict__ row = L.x + (size_t)L.feats[f] * L.ld;
    const uint32_t i0 = blockIdx.x * kChunk, i1 = min(L.m, i0 + kChunk);
    for (uint32_t base = i0; base < i1; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const int nid = i < i1 ? L.node_of[i] : -1;
        int uniform = 0;
        __match_all_sync(0xffffffffu, nid, &uniform);
        if (uniform && nid < 0) continue;
        uint32_t b = 0;
        int y = 0;
        bool positive = false;
        if (nid >= 0) {
            const uint32_t p = L.samp_pos[i];
            const double v = (double)__ldg(row + p);
            const double lo = (double)fdro(L.fmin[(size_t)nid * L.F + f]);
            const double range = __dsub_rn((double)fdro(L.fmax[(size_t)nid * L.F + f]), lo);
            while (b + 1 < k) {
                const double pos = __dadd_rn(__dmul_rn((double)(b + 1) / (double)k, range), lo);
                if (v < pos) break;
                ++b;
            }
            const float g = __ldg(L.gain + p);
            y = (int)__double2ll_rn((double)g * kGainScale);  // |label| <= 16: |y| <= 2^16
            positive = g > 0.0f;
        }
        if (SMEM && uniform && k <= 8) {
            // the whole warp sits in one node: per bucket, reduce in registers, one atomic each
            const long long sq64 = (long long)y * (long long)y;  // up to 2^32: reduced in two 16-bit halves
            for (uint32_t bb = 0; bb < k; ++bb) {
                const unsigned in = __ballot_sync(0xffffffffu, b == bb);
                if (!in) continue;
                const bool me = b == bb;
                const int cnt = __popc(in);
                const int npos = __popc(__ballot_sync(0xffffffffu, me && positive));
                const int ys = __reduce_add_sync(0xffffffffu, me ? y : 0);
                const unsigned qlo = __reduce_add_sync(0xffffffffu, me ? (unsigned)(sq64 & 0xffff) : 0u);
                const unsigned qhi = __reduce_add_sync(0xffffffffu, me ? (unsigned)(sq64 >> 16) : 0u);
                if ((threadIdx.x & 31) == 0) {
                    unsigned *cell = scell + 4 * ((size_t)nid * k + bb);
                    atomicAdd(cell + 0, (unsigned)cnt | ((unsigned)npos << 16));
                    atomicAdd(cell + 1, (unsigned)ys);
                    atomicAdd(cell + 2, qlo);
                    atomicAdd(cell + 3, qhi);
                }
            }
            continue;
        }
        if (nid < 0) continue;
        const long long yl = (long long)y;
        const size_t cell = (size_t)nid * k + b;
        if (SMEM) {
            unsigned *sc = scell + 4 * cell;
            const unsigned long long sq = (unsigned long long)(yl * yl);
            atomicAdd(sc + 0, positive ? 0x10001u : 1u);
            atomicAdd(sc + 1, (unsigned)y);
            atomicAdd(sc + 2, (unsigned)(sq & 0xffffu));
            atomicAdd(sc + 3, (unsigned)(sq >> 16));
        } else {
            const size_t gc = ((size_t)nid * L.F + f) * k + b;
            atomicAdd(&L.b_n[gc], 1u);
            if (positive) atomicAdd(&L.b_pos[gc], 1u);
            atomicAdd((unsigned long long *)&L.b_sum[gc], (unsigned long long)yl);
            atomicAdd((unsigned long long *)&L.b_sq[gc], (unsigned long long)(yl * yl));
        }
    }
    if (SMEM) {
        __syncthreads();
        for (size_t c = threadIdx.x; c < cells; c += blockDim.x) {
            const unsigned *sc = scell + 4 * c;
            const unsigned n_c = sc[0] & 0xffffu, pos_c = sc[0] >> 16;
            if (!n_c) continue;
            const size_t nid = c / k, b = c % k;
            const size_t gc = (nid * L.F + f) * k + b;
            atomicAdd(&L.b_n[gc], n_c);
            if (pos_c) atomicAdd(&L.b_pos[gc], pos_c);
            atomicAdd((unsigned long long *)&L.b_sum[gc], (unsigned long long)(long long)(int)sc[1]);
            atomicAdd((unsigned long long *)&L.b_sq[gc], ((unsigned long long)sc[3] << 16) + (unsigned long long)sc[2]);
        }
    }
}

// apply the host's decisions: every instance of a split node moves to its child
__global__ void rf_partition_kernel(const float *__restrict__ x, size_t ld, const uint32_t *__restrict__ samp_pos,
                                    int *__restrict__ node_of, uint32_t m, const uint32_t *__restrict__ fid,
                                    const double *__restrict__ split, const int *__restrict__ left,
                                    const int *__restrict__ right) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int nid = node_of[i];
    if (nid < 0) return;
    const uint32_t f = fid[nid];
    if (f == 0xffffffffu) {
        node_of[i] = -1;
        return;
    }
    const double v = (double)__ldg(x + (size_t)f * ld + samp_pos[i]);
    node_of[i] = v < split[nid] ? left[nid] : right[nid];  // lhs = values below the threshold
}

}  // namespace

struct fr_dev_rf {
    fr_dev_dataset *ds = nullptr;
    cudaStream_t stream = nullptr;
    uint32_t m = 0, F = 0;
    DevBuf<uint32_t> samp_pos, feats, t_fid;
    DevBuf<int> node_of, t_left, t_right;
    DevBuf<double> t_split;
    DevBuf<unsigned char> stats;  // one blob per level, cleared and read back in one go
    ~fr_dev_rf() {
        if (stream) cudaStreamDestroy(stream);
    }
};

extern "C" {

int fr_dev_rf_create(fr_dev_dataset *ds, fr_dev_rf **out) {
    if (!ds || !out) return fail("fr_dev_rf_create: NULL argument");
    CU(cudaSetDevice(ds->device));
    std::unique_ptr<fr_dev_rf> rf(new fr_dev_rf());
    rf->ds = ds;
    CU(cudaStreamCreateWithFlags(&rf->stream, cudaStreamNonBlocking));
    *out = rf.release();
    return 0;
}

void fr_dev_rf_destroy(fr_dev_rf *rf) {
    if (!rf) return;
    cudaSetDevice(rf->ds->device);
    delete rf;
}

int fr_dev_rf_begin_tree(fr_dev_rf *rf, const uint32_t *instances, size_t m, const uint32_t *features,
                         size_t n_features) {
    if (!rf || (!instances && m) || !features || n_features == 0) return fail("fr_dev_rf_begin_tree: bad argument");
    fr_dev_dataset *ds = rf->ds;
    CU(cudaSetDevice(ds->device));
    std::vector<uint32_t> pos(m);
    for (size_t i = 0; i < m; ++i) {
        if (instances[i] >= ds->n) return fail("fr_dev_rf_begin_tree: instance id out of range");
        pos[i] = ds->pos_of_inst[instances[i]];
    }
    for (size_t a = 0; a < n_features; ++a)
        if (features[a] >= ds->d) return fail("fr_dev_rf_begin_tree: feature id out of range");
    rf->m = (uint32_t)m;
    rf->F = (uint32_t)n_features;
    CU(rf->samp_pos.ensure(std::max<size_t>(m, 1)));
    CU(rf->node_of.ensure(std::max<size_t>(m, 1)));
    CU(rf->feats.ensure(n_features));
    if (m) CU(cudaMemcpyAsync(rf->samp_pos.p, pos.data(), sizeof(uint32_t) * m, cudaMemcpyHostToDevice, rf->stream));
    CU(cudaMemcpyAsync(rf->feats.p, features, sizeof(uint32_t) * n_features, cudaMemcpyHostToDevice, rf->stream));
    CU(cudaMemsetAsync(rf->node_of.p, 0, sizeof(int) * std::max<size_t>(m, 1), rf->stream));  // everyone in the root
    CU(cudaStreamSynchronize(rf->stream));  // `pos` is a local
    return 0;
}

int fr_dev_rf_level_stats(fr_dev_rf *rf, uint32_t n_active, uint32_t k, uint64_t *node_n, int64_t *node_sum,
                          float *gmin, float *gmax, float *fmin, float *fmax, uint32_t *b_n, uint32_t *b_pos,
                          int64_t *b_sum, int64_t *b_sq, uint32_t *f_present) {
    if (!rf || n_active == 0 || k < 2) return fail("fr_dev_rf_level_stats: bad argument");
    fr_dev_dataset *ds = rf->ds;
    CU(cudaSetDevice(ds->device));
    cudaStream_t s = rf->stream;
    const size_t nf = (size_t)n_active * rf->F, cells = nf * k;
    // blob: [node_n u64][node_sum i64][b_sum i64][b_sq i64][fmin i32][fmax i32][gmin][gmax][b_n u32][b_pos u32]
    const size_t o_node_n = 0, o_node_sum = o_node_n + 8 * (size_t)n_active, o_bsum = o_node_sum + 8 * (size_t)n_active,
                 o_bsq = o_bsum + 8 * cells, o_fmin = o_bsq + 8 * cells, o_fmax = o_fmin + 4 * nf,
                 o_gmin = o_fmax + 4 * nf, o_gmax = o_gmin + 4 * (size_t)n_active, o_bn = o_gmax + 4 * (size_t)n_active,
                 o_bpos = o_bn + 4 * cells, o_fp = o_bpos + 4 * cells, total = o_fp + 4 * nf;
    CU(rf->stats.ensure(total));
    unsigned char *b = rf->stats.p;
    CU(cudaMemsetAsync(b, 0, total, s));
    // min slots start at INT_MAX, max slots at INT_MIN
    {
        std::vector<int> init(2 * nf + 2 * (size_t)n_active);
        std::fill(init.begin(), init.begin() + nf, INT_MAX);
        std::fill(init.begin() + nf, init.begin() + 2 * nf, INT_MIN);
        std::fill(init.begin() + 2 * nf, init.begin() + 2 * nf + n_active, INT_MAX);
        std::fill(init.begin() + 2 * nf + n_active, init.end(), INT_MIN);
        CU(cudaMemcpyAsync(b + o_fmin, init.data(), sizeof(int) * init.size(), cudaMemcpyHostToDevice, s));
        CU(cudaStreamSynchronize(s));
    }
    RfLevel L;
    L.x = ds->x.p;
    L.ld = ds->ld;
    L.gain = ds->gain.p;
    L.len_pos = ds->len_pos.n ? ds->len_pos.p : nullptr;
    L.present_bits = ds->present_pos.n ? ds->present_pos.p : nullptr;
    L.present_words = ds->present_words;
    L.f_present = (unsigned *)(b + o_fp);
    L.samp_pos = rf->samp_pos.p;
    L.node_of = rf->node_of.p;
    L.m = rf->m;
    L.feats = rf->feats.p;
    L.F = rf->F;
    L.n_active = n_active;
    L.k = k;
    L.node_n = (unsigned long long *)(b + o_node_n);
    L.node_sum = (long long *)(b + o_node_sum);
    L.b_sum = (long long *)(b + o_bsum);
    L.b_sq = (long long *)(b + o_bsq);
    L.fmin = (int *)(b + o_fmin);
    L.fmax = (int *)(b + o_fmax);
    L.gmin = (int *)(b + o_gmin);
    L.gmax = (int *)(b + o_gmax);
    L.b_n = (unsigned *)(b + o_bn);
    L.b_pos = (unsigned *)(b + o_bpos);
    const dim3 grid((rf->m + kChunk - 1) / kChunk, rf->F);
    if (rf->m > 0) {
        const size_t smem1 = sizeof(int) * 4 * (size_t)n_active + sizeof(unsigned) * ((size_t)n_active + 1) +
                             sizeof(long long) * (size_t)n_active + sizeof(unsigned) * (size_t)n_active + 16;
        if (smem1 > 200 * 1024) return fail("fr_dev_rf_level_stats: too many active nodes");
        if (smem1 > 48 * 1024)
            CU(cudaFuncSetAttribute(rf_minmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        rf_minmax_kernel<<<grid, kRfThreads, smem1, s>>>(L);
        LAUNCHED();
        CU(cudaGetLastError());
        const size_t smem2 = (size_t)n_active * k * 16;
        if (smem2 <= 96 * 1024) {
            if (smem2 > 48 * 1024)
                CU(cudaFuncSetAttribute(rf_bucket_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            rf_bucket_kernel<true><<<grid, kRfThreads, smem2, s>>>(L);
        } else {
            rf_bucket_kernel<false><<<grid, kRfThreads, 0, s>>>(L);
        }
        LAUNCHED();
        CU(cudaGetLastError());
    }
    std::vector<unsigned char> host(total);
    CU(cudaMemcpyAsync(host.data(), b, total, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    memcpy(node_n, host.data() + o_node_n, 8 * (size_t)n_active);
    memcpy(node_sum, host.data() + o_node_sum, 8 * (size_t)n_active);
    memcpy(b_sum, host.data() + o_bsum, 8 * cells);
    memcpy(b_sq, host.data() + o_bsq, 8 * cells);
    memcpy(b_n, host.data() + o_bn, 4 * cells);
    memcpy(b_pos, host.data() + o_bpos, 4 * cells);
    if (f_present) {
        if (L.len_pos || L.present_bits) {
            memcpy(f_present, host.data() + o_fp, 4 * nf);
        } else {  // nothing is missing: every instance of the node carries every feature
            for (size_t i = 0; i < nf; ++i) f_present[i] = (uint32_t)node_n[i / rf->F];
        }
    }
    auto unord = [](int i) {
        const int j = i >= 0 ? i : i ^ 0x7fffffff;
        float f;
        memcpy(&f, &j, 4);
        return f;
    };
    const int *hfmin = (const int *)(host.data() + o_fmin), *hfmax = (const int *)(host.data() + o_fmax);
    const int *hgmin = (const int *)(host.data() + o_gmin), *hgmax = (const int *)(host.data() + o_gmax);
    for (size_t i = 0; i < nf; ++i) {
        fmin[i] = unord(hfmin[i]);
        fmax[i] = unord(hfmax[i]);
    }
    for (uint32_t a = 0; a < n_active; ++a) {
        gmin[a] = unord(hgmin[a]);
        gmax[a] = unord(hgmax[a]);
    }
    return 0;
}

int fr_dev_rf_partition(fr_dev_rf *rf, uint32_t n_active, const uint32_t *fid, const double *split,
                        const int32_t *left, const int32_t *right) {
    if (!rf || !fid || !split || !left || !right) return fail("fr_dev_rf_partition: NULL argument");
    fr_dev_dataset *ds = rf->ds;
    CU(cudaSetDevice(ds->device));
    cudaStream_t s = rf->stream;
    if (rf->m == 0 || n_active == 0) return 0;
    CU(rf->t_fid.ensure(n_active));
    CU(rf->t_split.ensure(n_active));
    CU(rf->t_left.ensure(n_active));
    CU(rf->t_right.ensure(n_active));
    CU(cudaMemcpyAsync(rf->t_fid.p, fid, sizeof(uint32_t) * n_active, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(rf->t_split.p, split, sizeof(double) * n_active, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(rf->t_left.p, left, sizeof(int) * n_active, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(rf->t_right.p, right, sizeof(int) * n_active, cudaMemcpyHostToDevice, s));
    rf_partition_kernel<<<(rf->m + 255) / 256, 256, 0, s>>>(ds->x.p, ds->ld, rf->samp_pos.p, rf->node_of.p, rf->m,
                                                           rf->t_fid.p, rf->t_split.p, rf->t_left.p, rf->t_right.p);
    LAUNCHED();
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));
    return 0;
}

}  // extern "C"
