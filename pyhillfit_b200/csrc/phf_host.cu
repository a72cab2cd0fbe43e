// Host-buffer entry points: the calls a numpy-level user of the reference would make.  Copy in, run the
// fused sampler in segments whose sample write-back (D2H) overlaps the next segment's kernel, copies out.
#include <algorithm>
#include <vector>

#include "phf_common.cuh"

using namespace phf;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
};

// Per-device workspace reused across calls (a sampler is called once per segment of a long run).
struct Workspace {
    DevBuf state, dsid, temp, datasets, groups, samples[2];
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    bool init = false;
};

Workspace g_ws[16][2];  // [device][model-1]: the two models may be driven from two host threads
Workspace g_ws_hier[16][8];  // [device][min(n_expts, 7)]: launches of different dimension may be driven concurrently

// The entry points run on `device` and leave the calling thread's current device as they found it.
struct DeviceGuard {
    int prev = -1;
    cudaError_t enter(int device)
    {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) { prev = -1; return e; }
        return prev == device ? cudaSuccess : cudaSetDevice(device);
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int ensure_streams(Workspace &w)
{
    if (w.init) return PHF_OK;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&w.compute, cudaStreamNonBlocking))) return set_cuda_error(e, "stream");
    if ((e = cudaStreamCreateWithFlags(&w.copy, cudaStreamNonBlocking))) return set_cuda_error(e, "stream");
    for (int i = 0; i < 2; ++i) {
        cudaEventCreateWithFlags(&w.done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&w.copied[i], cudaEventDisableTiming);
    }
    w.init = true;
    return PHF_OK;
}

// The segmented run shared by the two host entry points: cfg->n_iters iterations in segments on whole-row boundaries,
// segment k's rows copied to the host (copy stream) while segment k+1 computes.  `launch(c, dev_samples)` enqueues one
// segment on w.compute.  Device state is w.state (already enqueued H2D on w.compute); copied back at the end.
template <class Launch>
int run_segments(Workspace &w, const phf_am_config *cfg, int64_t n_chains, int d, double *state, double *samples,
                 int32_t n_segments, const char *what, Launch launch)
{
    const int nf = PHF_STATE_SIZE(d);
    cudaError_t e;
    const uint32_t total = cfg->n_iters;
    uint32_t seg_iters = (total + n_segments - 1) / n_segments;
    seg_iters = std::max<uint32_t>(cfg->thinning, (seg_iters + cfg->thinning - 1) / cfg->thinning * cfg->thinning);
    const uint32_t seg_rows_cap = seg_iters / cfg->thinning + 1;
    const size_t row_bytes = (size_t)(d + 1) * sizeof(double);
    if (samples)
        for (int i = 0; i < 2; ++i)
            if ((e = w.samples[i].ensure((size_t)n_chains * seg_rows_cap * row_bytes)))
                return set_cuda_error(e, "cudaMalloc(samples)");
    cudaStream_t cs = w.compute;
    uint32_t done_iters = 0, rows_done = 0;
    int seg = 0, rc = PHF_OK;
    while (done_iters < total) {
        const int b = seg & 1;
        phf_am_config c = *cfg;
        c.t0 = cfg->t0 + done_iters;
        c.n_iters = std::min(seg_iters, total - done_iters);
        c.rows_capacity = seg_rows_cap;
        const uint32_t rows = (c.t0 + c.n_iters) / c.thinning - c.t0 / c.thinning;
        if (samples && seg >= 2) cudaStreamWaitEvent(cs, w.copied[b], 0);  // buffer b must have been drained
        rc = launch(c, samples ? (double *)w.samples[b].p : nullptr);
        if (rc != PHF_OK) break;
        if (samples && rows > 0) {
            cudaEventRecord(w.done[b], cs);
            cudaStreamWaitEvent(w.copy, w.done[b], 0);
            if (cfg->sample_layout == PHF_SAMPLES_ROW_MAJOR) {
                // device [seg_rows_cap][chain][d+1] -> host [rows_capacity][chain][d+1] at row rows_done: contiguous
                e = cudaMemcpyAsync(samples + (size_t)rows_done * n_chains * (d + 1), w.samples[b].p,
                                    (size_t)rows * n_chains * row_bytes, cudaMemcpyDeviceToHost, w.copy);
            } else {
                // device [chain][seg_rows_cap][d+1] -> host [chain][rows_capacity][d+1] at row offset rows_done
                e = cudaMemcpy2DAsync(samples + (size_t)rows_done * (d + 1), (size_t)cfg->rows_capacity * row_bytes,
                                      w.samples[b].p, (size_t)seg_rows_cap * row_bytes, (size_t)rows * row_bytes,
                                      (size_t)n_chains, cudaMemcpyDeviceToHost, w.copy);
            }
            if (e) { rc = set_cuda_error(e, "cudaMemcpyAsync(samples)"); break; }
            cudaEventRecord(w.copied[b], w.copy);
        }
        done_iters += c.n_iters;
        rows_done += rows;
        ++seg;
    }
    if (rc == PHF_OK)
        cudaMemcpyAsync(state, w.state.p, (size_t)n_chains * nf * sizeof(double), cudaMemcpyDeviceToHost, cs);
    cudaError_t e1 = cudaStreamSynchronize(cs), e2 = cudaStreamSynchronize(w.copy);
    if (rc != PHF_OK) return rc;
    if (e1) return set_cuda_error(e1, what);
    if (e2) return set_cuda_error(e2, what);
    return PHF_OK;
}

}  // namespace

extern "C" int phf_am_single_run_host(const phf_am_config *cfg, int64_t n_chains, double *state,
                                      const int32_t *dataset_id, const double *temperature, int32_t n_datasets,
                                      const phf_dataset *datasets, int32_t n_groups, const phf_dose_group *groups,
                                      double *samples, int32_t n_segments, int32_t device)
{
    if (!cfg) return set_error(PHF_EINVAL, "phf_am_single_run_host: cfg is NULL");
    if (cfg->model != 1 && cfg->model != 2) return set_error(PHF_EINVAL, "cfg.model must be 1 or 2");
    if (cfg->thinning == 0) return set_error(PHF_EINVAL, "cfg.thinning must be >= 1");
    if (cfg->sample_layout != PHF_SAMPLES_CHAIN_MAJOR && cfg->sample_layout != PHF_SAMPLES_ROW_MAJOR)
        return set_error(PHF_EINVAL, "cfg.sample_layout must be PHF_SAMPLES_CHAIN_MAJOR or PHF_SAMPLES_ROW_MAJOR");
    if (device < 0 || device >= 16) return set_error(PHF_EINVAL, "device index outside 0..15");
    if (n_chains <= 0 || n_datasets <= 0 || n_groups <= 0 || !state || !dataset_id || !temperature || !datasets ||
        !groups)
        return set_error(PHF_EINVAL, "phf_am_single_run_host: empty or null input");
    if (n_segments < 1) n_segments = 1;
    const int d = cfg->model == 1 ? 2 : 3, nf = PHF_STATE_SIZE(d);
    cudaError_t e;
    DeviceGuard guard;
    if ((e = guard.enter(device))) return set_cuda_error(e, "cudaSetDevice");
    Workspace &w = g_ws[device][cfg->model - 1];
    if (int rc = ensure_streams(w)) return rc;
    const uint32_t rows_total = (cfg->t0 + cfg->n_iters) / cfg->thinning - cfg->t0 / cfg->thinning;
    if (samples && rows_total > cfg->rows_capacity)
        return set_error(PHF_EINVAL, "cfg.rows_capacity is smaller than the rows this call produces");

    if ((e = w.state.ensure((size_t)n_chains * nf * sizeof(double))) ||
        (e = w.dsid.ensure((size_t)n_chains * sizeof(int32_t))) ||
        (e = w.temp.ensure((size_t)n_chains * sizeof(double))) ||
        (e = w.datasets.ensure((size_t)n_datasets * sizeof(phf_dataset))) ||
        (e = w.groups.ensure((size_t)n_groups * sizeof(phf_dose_group))))
        return set_cuda_error(e, "cudaMalloc");

    cudaStream_t cs = w.compute;
    cudaMemcpyAsync(w.state.p, state, (size_t)n_chains * nf * sizeof(double), cudaMemcpyHostToDevice, cs);
    cudaMemcpyAsync(w.dsid.p, dataset_id, (size_t)n_chains * sizeof(int32_t), cudaMemcpyHostToDevice, cs);
    cudaMemcpyAsync(w.temp.p, temperature, (size_t)n_chains * sizeof(double), cudaMemcpyHostToDevice, cs);
    cudaMemcpyAsync(w.datasets.p, datasets, (size_t)n_datasets * sizeof(phf_dataset), cudaMemcpyHostToDevice, cs);
    cudaMemcpyAsync(w.groups.p, groups, (size_t)n_groups * sizeof(phf_dose_group), cudaMemcpyHostToDevice, cs);

    return run_segments(w, cfg, n_chains, d, state, samples, n_segments, "phf_am_single_run_host",
                        [&](const phf_am_config &c, double *dev_samples) {
                            return phf_am_single_run(&c, n_chains, (double *)w.state.p, (const int32_t *)w.dsid.p,
                                                     (const double *)w.temp.p, (const phf_dataset *)w.datasets.p,
                                                     (const phf_dose_group *)w.groups.p, dev_samples, cs);
                        });
}

// Hierarchical counterpart (all chains of a call share n_expts, like phf_am_hier_run).
extern "C" int phf_am_hier_run_host(const phf_am_config *cfg, int32_t n_expts, int64_t n_chains, double *state,
                                    const int32_t *dataset_id, int32_t n_datasets, const phf_hier_dataset *datasets,
                                    int32_t n_points, const phf_hier_point *points, const phf_hier_priors *priors,
                                    double *samples, int32_t n_segments, int32_t device)
{
    if (!cfg || !priors) return set_error(PHF_EINVAL, "phf_am_hier_run_host: cfg/priors is NULL");
    if (n_expts < 1 || n_expts > PHF_HIER_BIG_MAX_EXPTS) return set_error(PHF_ENOTSUP, "n_expts outside 1..128");
    if (cfg->thinning == 0) return set_error(PHF_EINVAL, "cfg.thinning must be >= 1");
    if (cfg->sample_layout != PHF_SAMPLES_CHAIN_MAJOR && cfg->sample_layout != PHF_SAMPLES_ROW_MAJOR)
        return set_error(PHF_EINVAL, "cfg.sample_layout must be PHF_SAMPLES_CHAIN_MAJOR or PHF_SAMPLES_ROW_MAJOR");
    if (device < 0 || device >= 16) return set_error(PHF_EINVAL, "device index outside 0..15");
    if (n_chains <= 0 || n_datasets <= 0 || n_points <= 0 || !state || !dataset_id || !datasets || !points)
        return set_error(PHF_EINVAL, "phf_am_hier_run_host: empty or null input");
    if (n_segments < 1) n_segments = 1;
    const int d = 5 + 2 * n_expts, nf = PHF_STATE_SIZE(d);
    cudaError_t e;
    DeviceGuard guard;
    if ((e = guard.enter(device))) return set_cuda_error(e, "cudaSetDevice");
    Workspace &w = g_ws_hier[device][n_expts < 7 ? n_expts : 7];
    if (int rc = ensure_streams(w)) return rc;
    const uint32_t rows_total = (cfg->t0 + cfg->n_iters) / cfg->thinning - cfg->t0 / cfg->thinning;
    if (samples && rows_total > cfg->rows_capacity)
        return set_error(PHF_EINVAL, "cfg.rows_capacity is smaller than the rows this call produces");

    // (w.datasets / w.groups hold the hierarchical datasets / points here)
    if ((e = w.state.ensure((size_t)n_chains * nf * sizeof(double))) ||
        (e = w.dsid.ensure((size_t)n_chains * sizeof(int32_t))) ||
        (e = w.datasets.ensure((size_t)n_datasets * sizeof(phf_hier_dataset))) ||
        (e = w.groups.ensure((size_t)n_points * sizeof(phf_hier_point))))
        return set_cuda_error(e, "cudaMalloc");

    cudaStream_t cs = w.compute;
    cudaMemcpyAsync(w.state.p, state, (size_t)n_chains * nf * sizeof(double), cudaMemcpyHostToDevice, cs);
    cudaMemcpyAsync(w.dsid.p, dataset_id, (size_t)n_chains * sizeof(int32_t), cudaMemcpyHostToDevice, cs);
    cudaMemcpyAsync(w.datasets.p, datasets, (size_t)n_datasets * sizeof(phf_hier_dataset), cudaMemcpyHostToDevice, cs);
    cudaMemcpyAsync(w.groups.p, points, (size_t)n_points * sizeof(phf_hier_point), cudaMemcpyHostToDevice, cs);

    return run_segments(w, cfg, n_chains, d, state, samples, n_segments, "phf_am_hier_run_host",
                        [&](const phf_am_config &c, double *dev_samples) {
                            return phf_am_hier_run(&c, n_expts, n_chains, (double *)w.state.p,
                                                   (const int32_t *)w.dsid.p, (const phf_hier_dataset *)w.datasets.p,
                                                   (const phf_hier_point *)w.groups.p, priors, dev_samples, cs);
                        });
}
