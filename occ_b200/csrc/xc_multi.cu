// xc_multi.cu — the XC build on several GPUs of one box, driven from ONE host process behind the C ABI
// (occ_xc_multi_*, include/occ_xc_b200.h).  This is the layout the reference needs: its SCF is a single process
// (include/occ/qm/scf_impl.h:539 calls compute_fock on one thread) and the leaf loop it parallelises with
// tbb::parallel_for over thread-local K copies (include/occ/dft/dft.h:221-232,303-371), followed by a serial
// reduction of those copies (dft.h:375-380), becomes
//   one handle (= one leaf shard, occ_xc_partition_leaves) per GPU, each driven by its own host worker thread;
//   inputs  : every GPU pulls 1/N of D / C from the host over ITS OWN PCIe link, then gathers the other slices from
//             its peers over NVLink (cudaMemcpyPeerAsync) — the host buffer crosses PCIe once in total;
//   outputs : GPU r owns column panel r of K.  One kernel sums that panel over all partial K's in rank order through
//             peer loads (NVSwitch; fixed order = deterministic) and GPU r writes its panel to the host matrix.
// No NCCL in this path: with all handles in one address space the reduce-scatter is a peer-memory kernel.
#include "xc_handle.h"
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>

namespace {

constexpr int MAX_DEV = 16;

struct PeerPtrs {
  const double *p[MAX_DEV];
  int n;
};

// out[i] = sum_r part[r][i] in rank order, i in [begin, end)
__global__ void k_reduce_panel(PeerPtrs parts, double *__restrict__ out, size_t begin, size_t end) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = begin + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < end; i += stride) {
    double s = parts.p[0][i];
    for (int r = 1; r < parts.n; ++r) s += parts.p[r][i];
    out[i] = s;
  }
}

// One persistent host thread per device: issuing ~70 launches per build takes a few hundred microseconds, which in
// series over 8 devices would delay the last GPU by milliseconds of a 20 ms build.
class Worker {
public:
  Worker() : m_thread([this] { loop(); }) {}
  ~Worker() {
    {
      std::lock_guard<std::mutex> g(m_mu);
      m_stop = true;
    }
    m_cv.notify_all();
    m_thread.join();
  }
  void submit(std::function<int()> job) {
    {
      std::lock_guard<std::mutex> g(m_mu);
      m_job = std::move(job);
      m_has_job = true;
      m_done = false;
    }
    m_cv.notify_all();
  }
  int wait(std::string *err) {
    std::unique_lock<std::mutex> g(m_mu);
    m_cv.wait(g, [this] { return m_done; });
    if (m_rc && err) *err = m_err;
    return m_rc;
  }

private:
  void loop() {
    for (;;) {
      std::function<int()> job;
      {
        std::unique_lock<std::mutex> g(m_mu);
        m_cv.wait(g, [this] { return m_has_job || m_stop; });
        if (m_stop) return;
        job = std::move(m_job);
        m_has_job = false;
      }
      const int rc = job();
      std::string err = rc ? occ_xc_last_error() : ""; // thread-local message of THIS worker
      {
        std::lock_guard<std::mutex> g(m_mu);
        m_rc = rc;
        m_err = err;
        m_done = true;
      }
      m_cv.notify_all();
    }
  }
  std::mutex m_mu;
  std::condition_variable m_cv;
  std::function<int()> m_job;
  bool m_has_job = false, m_done = true, m_stop = false;
  int m_rc = 0;
  std::string m_err;
  std::thread m_thread;
};

} // namespace

struct occ_xc_multi {
  int n = 0;
  int device[MAX_DEV];
  occ_xc_handle *h[MAX_DEV];
  Worker *worker[MAX_DEV];
  cudaEvent_t ev_in[MAX_DEV];    // this rank's slice of the inputs is on its device
  cudaEvent_t ev_built[MAX_DEV]; // this rank's partial K is complete
  double *d_red[MAX_DEV];        // reduced panel (full-size buffer, only the rank's panel is written)
  double *h_scalars = nullptr;   // pinned, 3 per rank
  size_t ksz = 0, csz = 0;
  double last_ms[MAX_DEV];
  // post-SCF VV10 over the shards (allocated on first use)
  cudaEvent_t ev_vv[MAX_DEV];    // this rank's coefficient / contraction stage is complete
  double *d_vv2[MAX_DEV];        // merged per-point arrays (X Y Z rw2 om kap) and summed F U W: 9 n doubles
  double *h_grad = nullptr;      // pinned, 3 natoms per rank
};

#define MULTI_TRY(expr)                                                                                        \
  do {                                                                                                         \
    cudaError_t e__ = (expr);                                                                                  \
    if (e__ != cudaSuccess)                                                                                    \
      return occ_xc_set_error((std::string("CUDA error: ") + cudaGetErrorString(e__) + " in " #expr).c_str());  \
  } while (0)

// element range [lo, hi) of rank r when `total` doubles are cut into n contiguous slices on even boundaries
static void slice(size_t total, int n, int r, size_t *lo, size_t *hi) {
  const size_t per = ((total + n - 1) / n + 1) & ~size_t(1);
  *lo = std::min(total, per * r);
  *hi = std::min(total, per * (r + 1));
}
// column panel of rank r: whole columns of the (ld x ncols) matrix, so that a panel is one contiguous chunk
static void panel(size_t ld, size_t ncols, int n, int r, size_t *lo, size_t *hi) {
  const size_t per = (ncols + n - 1) / n;
  *lo = std::min(ncols, per * r) * ld;
  *hi = std::min(ncols, per * (r + 1)) * ld;
}

extern "C" {

void occ_xc_multi_destroy(occ_xc_multi *m) {
  if (!m) return;
  for (int r = 0; r < m->n; ++r) {
    delete m->worker[r];
    if (m->h[r]) {
      cudaSetDevice(m->device[r]);
      if (m->ev_in[r]) cudaEventDestroy(m->ev_in[r]);
      if (m->ev_built[r]) cudaEventDestroy(m->ev_built[r]);
      if (m->d_red[r]) cudaFree(m->d_red[r]);
      if (m->ev_vv[r]) cudaEventDestroy(m->ev_vv[r]);
      if (m->d_vv2[r]) cudaFree(m->d_vv2[r]);
      occ_xc_destroy(m->h[r]);
    }
  }
  if (m->h_scalars) cudaFreeHost(m->h_scalars);
  if (m->h_grad) cudaFreeHost(m->h_grad);
  delete m;
}

int occ_xc_multi_create(const occ_xc_basis_t *basis, size_t npts, const double *pts_xyz, const double *weights,
                        size_t nleaves, const occ_xc_leaf_t *leaves, const double *shell_extents, int nfunc,
                        const int *func_ids, int nspin, double density_threshold, double screening_threshold,
                        const occ_xc_options_t *options, int ndev, const int *devices, occ_xc_multi **out) {
  if (!out) return occ_xc_set_error("occ_xc_multi_create: null argument");
  if (ndev < 1 || ndev > MAX_DEV) return occ_xc_set_error("occ_xc_multi_create: between 1 and 16 devices");
  int avail = 0;
  if (cudaGetDeviceCount(&avail) != cudaSuccess || avail == 0)
    return occ_xc_set_error("occ_xc_multi_create: no CUDA device available (this library has no CPU fallback)");
  occ_xc_multi *m = new occ_xc_multi();
  m->n = ndev;
  for (int r = 0; r < MAX_DEV; ++r) {
    m->h[r] = nullptr;
    m->worker[r] = nullptr;
    m->ev_in[r] = m->ev_built[r] = nullptr;
    m->d_red[r] = nullptr;
    m->ev_vv[r] = nullptr;
    m->d_vv2[r] = nullptr;
    m->last_ms[r] = 0.0;
  }
  for (int r = 0; r < ndev; ++r) {
    m->device[r] = devices ? devices[r] : r;
    if (m->device[r] < 0 || m->device[r] >= avail) {
      occ_xc_multi_destroy(m);
      return occ_xc_set_error("occ_xc_multi_create: device ordinal out of range");
    }
  }
  // peer access between every pair of distinct devices (the same ordinal may appear twice: two shards on one GPU)
  for (int a = 0; a < ndev; ++a)
    for (int b = 0; b < ndev; ++b) {
      if (m->device[a] == m->device[b]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, m->device[a], m->device[b]);
      if (!can) {
        occ_xc_multi_destroy(m);
        return occ_xc_set_error("occ_xc_multi_create: the devices cannot access each other's memory (no NVLink / PCIe peer "
                                "path); use one process per GPU with an NCCL all-reduce instead");
      }
      cudaSetDevice(m->device[a]);
      const cudaError_t e = cudaDeviceEnablePeerAccess(m->device[b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        occ_xc_multi_destroy(m);
        return occ_xc_set_error((std::string("occ_xc_multi_create: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)).c_str());
      }
      cudaGetLastError();
    }
  // one handle per rank, created concurrently (each create screens the whole grid on the host and uploads its shard)
  for (int r = 0; r < ndev; ++r) m->worker[r] = new Worker();
  for (int r = 0; r < ndev; ++r) {
    occ_xc_options_t opt;
    if (options) opt = *options;
    else occ_xc_options_init(&opt);
    opt.device = m->device[r];
    opt.rank = r;
    opt.world_size = ndev;
    m->worker[r]->submit([=]() -> int {
      return occ_xc_create(basis, npts, pts_xyz, weights, nleaves, leaves, shell_extents, nfunc, func_ids, nspin,
                           density_threshold, screening_threshold, &opt, &m->h[r]);
    });
  }
  std::string err;
  int rc = 0;
  for (int r = 0; r < ndev; ++r) rc |= m->worker[r]->wait(&err);
  if (rc) {
    occ_xc_multi_destroy(m);
    return occ_xc_set_error(err.c_str());
  }
  m->ksz = static_cast<size_t>(nspin) * m->h[0]->nbf * m->h[0]->nbf;
  m->csz = static_cast<size_t>(m->h[0]->nbf) * (m->h[0]->nocc[0] + m->h[0]->nocc[1]);
  for (int r = 0; r < ndev; ++r) {
    cudaSetDevice(m->device[r]);
    if (cudaEventCreateWithFlags(&m->ev_in[r], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_built[r], cudaEventDisableTiming) != cudaSuccess ||
        cudaMalloc(&m->d_red[r], m->ksz * sizeof(double)) != cudaSuccess) {
      occ_xc_multi_destroy(m);
      return occ_xc_set_error("occ_xc_multi_create: out of device memory");
    }
  }
  if (cudaHostAlloc(&m->h_scalars, 3 * MAX_DEV * sizeof(double), cudaHostAllocPortable) != cudaSuccess) {
    occ_xc_multi_destroy(m);
    return occ_xc_set_error("occ_xc_multi_create: cudaHostAlloc failed");
  }
  *out = m;
  return 0;
}

int occ_xc_multi_ndev(const occ_xc_multi *m) { return m ? m->n : 0; }

int occ_xc_multi_get_info(const occ_xc_multi *m, int rank, occ_xc_info_t *info) {
  if (!m || rank < 0 || rank >= m->n) return occ_xc_set_error("occ_xc_multi_get_info: bad argument");
  return occ_xc_get_info(m->h[rank], info);
}

// One compute_K_dft call over all devices; HOST buffers.  D2 may be NULL when C is given (D2 = C C^T on each device).
int occ_xc_multi_build(occ_xc_multi *m, const double *D2, const double *C, double *K, double *exc, double *n_alpha,
                       double *n_beta) {
  if (!m || !K || (!D2 && !C)) return occ_xc_set_error("occ_xc_multi_build: null argument");
  if (C && !m->h[0]->mo) return occ_xc_set_error("occ_xc_multi_build: the handles were created without occupied-orbital counts");
  const int n = m->n;
  // stage 1 (each worker): this rank's slice of D2 / C  host -> own device
  for (int r = 0; r < n; ++r) {
    m->worker[r]->submit([=]() -> int {
      occ_xc_handle *h = m->h[r];
      MULTI_TRY(cudaSetDevice(h->device));
      size_t lo, hi;
      if (D2) {
        slice(m->ksz, n, r, &lo, &hi);
        if (hi > lo) {
          MULTI_TRY(cudaMemcpyAsync(h->d_D + lo, D2 + lo, (hi - lo) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
          // the scales are applied to the rank's OWN slice before it is published: peers gather scaled data
          if (h->density_scale != 1.0) xcb::launch_scale(h->d_D + lo, h->d_D + lo, hi - lo, h->density_scale, h->stream);
        }
      }
      if (C) {
        slice(m->csz, n, r, &lo, &hi);
        if (hi > lo) {
          MULTI_TRY(cudaMemcpyAsync(h->d_C + lo, C + lo, (hi - lo) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
          if (h->orbital_scale != 1.0) xcb::launch_scale(h->d_C + lo, h->d_C + lo, hi - lo, h->orbital_scale, h->stream);
        }
      }
      MULTI_TRY(cudaEventRecord(m->ev_in[r], h->stream));
      return 0;
    });
  }
  std::string err;
  int rc = 0;
  for (int r = 0; r < n; ++r) rc |= m->worker[r]->wait(&err);
  if (rc) return occ_xc_set_error(err.c_str());
  // stage 2 (each worker): gather the other slices from the peers, scale, build, record
  for (int r = 0; r < n; ++r) {
    m->worker[r]->submit([=]() -> int {
      occ_xc_handle *h = m->h[r];
      MULTI_TRY(cudaSetDevice(h->device));
      for (int k = 1; k < n; ++k) {
        const int p = (r + k) % n; // staggered, so that the peers are not all read in the same order
        MULTI_TRY(cudaStreamWaitEvent(h->stream, m->ev_in[p], 0));
        size_t lo, hi;
        if (D2) {
          slice(m->ksz, n, p, &lo, &hi);
          if (hi > lo)
            MULTI_TRY(cudaMemcpyPeerAsync(h->d_D + lo, h->device, m->h[p]->d_D + lo, m->h[p]->device, (hi - lo) * sizeof(double), h->stream));
        }
        if (C) {
          slice(m->csz, n, p, &lo, &hi);
          if (hi > lo)
            MULTI_TRY(cudaMemcpyPeerAsync(h->d_C + lo, h->device, m->h[p]->d_C + lo, m->h[p]->device, (hi - lo) * sizeof(double), h->stream));
        }
      }
      if (xc_build_device_impl(h, D2 ? h->d_D : nullptr, C ? h->d_C : nullptr, h->d_K, h->d_scalars, h->stream)) return 1;
      MULTI_TRY(cudaMemcpyAsync(m->h_scalars + 3 * r, h->d_scalars, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      MULTI_TRY(cudaEventRecord(m->ev_built[r], h->stream));
      return 0;
    });
  }
  for (int r = 0; r < n; ++r) rc |= m->worker[r]->wait(&err);
  if (rc) return occ_xc_set_error(err.c_str());
  // stage 3 (each worker): reduce this rank's column panel over all partial K's (peer loads), write it to the host
  const size_t ld = static_cast<size_t>(m->h[0]->nspin) * m->h[0]->nbf, ncols = m->h[0]->nbf;
  for (int r = 0; r < n; ++r) {
    m->worker[r]->submit([=]() -> int {
      occ_xc_handle *h = m->h[r];
      MULTI_TRY(cudaSetDevice(h->device));
      size_t lo, hi;
      panel(ld, ncols, n, r, &lo, &hi);
      for (int p = 0; p < n; ++p)
        if (p != r) MULTI_TRY(cudaStreamWaitEvent(h->stream, m->ev_built[p], 0));
      if (hi > lo) {
        PeerPtrs parts;
        parts.n = n;
        for (int p = 0; p < n; ++p) parts.p[p] = m->h[p]->d_K;
        const unsigned grid = static_cast<unsigned>(std::min<size_t>((hi - lo + 255) / 256, 148 * 8));
        k_reduce_panel<<<grid, 256, 0, h->stream>>>(parts, m->d_red[r], lo, hi);
        MULTI_TRY(cudaMemcpyAsync(K + lo, m->d_red[r] + lo, (hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      }
      MULTI_TRY(cudaStreamSynchronize(h->stream));
      MULTI_TRY(cudaGetLastError());
      return 0;
    });
  }
  for (int r = 0; r < n; ++r) rc |= m->worker[r]->wait(&err);
  if (rc) return occ_xc_set_error(err.c_str());
  // a peer may still be reading this rank's partial K when the next build starts overwriting it: every stream has
  // been synchronised above, and stage 3 of rank r only completes after its reads, so all reads are done here.
  double s[3] = {0.0, 0.0, 0.0};
  for (int r = 0; r < n; ++r) // rank order: the reference's serial reduction over thread copies (dft.h:375-380)
    for (int k = 0; k < 3; ++k) s[k] += m->h_scalars[3 * r + k];
  if (exc) *exc = s[0];
  if (n_alpha) *n_alpha = s[1];
  if (n_beta) *n_beta = s[2];
  return 0;
}

// ---- post-SCF VV10 over the shards (NonLocalCorrelationFunctional::vv10 / vv10_gradient,
// src/dft/nonlocal_correlation.cpp:367-420, 392-522).  Every rank evaluates rho / grad rho on ITS leaves and publishes
// the per-point arrays of its points (reference order); each rank then gathers the whole grid's arrays from its peers,
// runs its round-robin share of the O(N^2) pair work, the partial F / U / W are summed through peer loads in rank order,
// and the Vxc contraction (or the gradient's potential term) runs on the rank's own leaves again.
static int multi_vv10(occ_xc_multi *m, const double *D2, const size_t *orig_index, const size_t *atom_offsets, double vv10_b,
                      double vv10_C, double *K, double *gradient, double *energy) {
  const int n = m->n;
  const bool grad = gradient != nullptr;
  const char *who = grad ? "occ_xc_multi_vv10_gradient" : "occ_xc_multi_vv10";
  for (int r = 0; r < n; ++r)
    if (vv10_check_handle(m->h[r], who, grad)) return 1;
  const size_t npts = m->h[0]->npts_total;
  const int natoms = m->h[0]->natoms;
  if (grad && !m->h_grad &&
      cudaHostAlloc(&m->h_grad, 3 * static_cast<size_t>(natoms) * MAX_DEV * sizeof(double), cudaHostAllocPortable) != cudaSuccess)
    return occ_xc_set_error("occ_xc_multi_vv10_gradient: cudaHostAlloc failed");
  std::string err;
  int rc = 0;
  // stage 1: density on the rank's own leaves, its points' arrays published in h->vv (0 = "not mine")
  for (int r = 0; r < n; ++r) {
    m->worker[r]->submit([=]() -> int {
      occ_xc_handle *h = m->h[r];
      MULTI_TRY(cudaSetDevice(h->device));
      if (!m->ev_vv[r]) MULTI_TRY(cudaEventCreateWithFlags(&m->ev_vv[r], cudaEventDisableTiming));
      if (!m->d_vv2[r]) MULTI_TRY(cudaMalloc(&m->d_vv2[r], 9 * npts * sizeof(double)));
      if (vv10_setup(h, r, n)) return 1;
      if (grad && vv10_setup_gradient(h, r, n, atom_offsets)) return 1;
      if (vv10_stage_density(h, D2, orig_index, vv10_b, vv10_C, 0.0, grad)) return 1;
      MULTI_TRY(cudaEventRecord(m->ev_in[r], h->stream));
      return 0;
    });
  }
  for (int r = 0; r < n; ++r) rc |= m->worker[r]->wait(&err);
  if (rc) return occ_xc_set_error(err.c_str());
  auto second_set = [=](int r) {
    xcb::VV10Params V = m->h[r]->vv;
    double *b = m->d_vv2[r];
    V.X = b, V.Y = b + npts, V.Z = b + 2 * npts, V.rw2 = b + 3 * npts, V.om = b + 4 * npts, V.kap = b + 5 * npts;
    V.F = b + 6 * npts, V.U = b + 7 * npts, V.W = b + 8 * npts;
    return V;
  };
  // stage 2: gather the whole grid's arrays, this rank's share of the pair work
  for (int r = 0; r < n; ++r) {
    m->worker[r]->submit([=]() -> int {
      occ_xc_handle *h = m->h[r];
      MULTI_TRY(cudaSetDevice(h->device));
      for (int p = 0; p < n; ++p)
        if (p != r) MULTI_TRY(cudaStreamWaitEvent(h->stream, m->ev_in[p], 0));
      xcb::VV10Peers src;
      src.n = n;
      for (int p = 0; p < n; ++p) src.base[p] = m->h[p]->vv.X;
      const xcb::VV10Params merged = second_set(r);
      xcb::launch_vv10_merge(src, merged, h->stream);
      if (vv10_stage_pairs(h, merged)) return 1;
      MULTI_TRY(cudaEventRecord(m->ev_built[r], h->stream));
      return 0;
    });
  }
  for (int r = 0; r < n; ++r) rc |= m->worker[r]->wait(&err);
  if (rc) return occ_xc_set_error(err.c_str());
  // stage 3: sum F / U / W over the ranks (rank order), then the contraction on the rank's own leaves
  for (int r = 0; r < n; ++r) {
    m->worker[r]->submit([=]() -> int {
      occ_xc_handle *h = m->h[r];
      MULTI_TRY(cudaSetDevice(h->device));
      for (int p = 0; p < n; ++p)
        if (p != r) MULTI_TRY(cudaStreamWaitEvent(h->stream, m->ev_built[p], 0));
      PeerPtrs parts;
      parts.n = n;
      for (int p = 0; p < n; ++p) parts.p[p] = m->h[p]->vv.F; // F, U, W are contiguous
      const xcb::VV10Params merged = second_set(r);
      const unsigned g = static_cast<unsigned>(std::min<size_t>((3 * npts + 255) / 256, 148 * 8));
      k_reduce_panel<<<g, 256, 0, h->stream>>>(parts, merged.F, 0, 3 * npts);
      if (grad) {
        if (vv10_stage_gradient(h, merged, merged)) return 1;
        MULTI_TRY(cudaMemcpyAsync(m->h_grad + 3 * static_cast<size_t>(natoms) * r, h->d_grad,
                                  3 * static_cast<size_t>(natoms) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      } else {
        if (vv10_stage_vxc(h, merged)) return 1;
      }
      MULTI_TRY(cudaMemcpyAsync(m->h_scalars + 3 * r, h->d_scalars, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      MULTI_TRY(cudaEventRecord(m->ev_vv[r], h->stream));
      return 0;
    });
  }
  for (int r = 0; r < n; ++r) rc |= m->worker[r]->wait(&err);
  if (rc) return occ_xc_set_error(err.c_str());
  // stage 4: column panels of K summed over the ranks and written to the host (as occ_xc_multi_build); sync
  const size_t ld = m->h[0]->nbf, ncols = m->h[0]->nbf;
  for (int r = 0; r < n; ++r) {
    m->worker[r]->submit([=]() -> int {
      occ_xc_handle *h = m->h[r];
      MULTI_TRY(cudaSetDevice(h->device));
      for (int p = 0; p < n; ++p)
        if (p != r) MULTI_TRY(cudaStreamWaitEvent(h->stream, m->ev_vv[p], 0));
      size_t lo, hi;
      panel(ld, ncols, n, r, &lo, &hi);
      if (!grad && hi > lo) {
        PeerPtrs parts;
        parts.n = n;
        for (int p = 0; p < n; ++p) parts.p[p] = m->h[p]->d_K;
        const unsigned g = static_cast<unsigned>(std::min<size_t>((hi - lo + 255) / 256, 148 * 8));
        k_reduce_panel<<<g, 256, 0, h->stream>>>(parts, m->d_red[r], lo, hi);
        MULTI_TRY(cudaMemcpyAsync(K + lo, m->d_red[r] + lo, (hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      }
      MULTI_TRY(cudaStreamSynchronize(h->stream));
      MULTI_TRY(cudaGetLastError());
      return 0;
    });
  }
  for (int r = 0; r < n; ++r) rc |= m->worker[r]->wait(&err);
  if (rc) return occ_xc_set_error(err.c_str());
  double e = 0.0;
  for (int r = 0; r < n; ++r) e += m->h_scalars[3 * r];
  if (energy) *energy = e;
  if (grad)
    for (int k = 0; k < 3 * natoms; ++k) {
      double g = 0.0;
      for (int r = 0; r < n; ++r) g += m->h_grad[3 * static_cast<size_t>(natoms) * r + k];
      gradient[k] = g;
    }
  return 0;
}

int occ_xc_multi_vv10(occ_xc_multi *m, const double *D2, const size_t *orig_index, double vv10_b, double vv10_C, double *K,
                      double *energy) {
  if (!m || !D2 || !K || !energy) return occ_xc_set_error("occ_xc_multi_vv10: null argument");
  return multi_vv10(m, D2, orig_index, nullptr, vv10_b, vv10_C, K, nullptr, energy);
}

int occ_xc_multi_vv10_gradient(occ_xc_multi *m, const double *D2, const size_t *orig_index, const size_t *atom_offsets,
                               double vv10_b, double vv10_C, double *gradient, double *energy) {
  if (!m || !D2 || !atom_offsets || !gradient) return occ_xc_set_error("occ_xc_multi_vv10_gradient: null argument");
  return multi_vv10(m, D2, orig_index, atom_offsets, vv10_b, vv10_C, nullptr, gradient, energy);
}

// XC nuclear gradient over the shards (DFT::compute_xc_gradient_impl, dft.h:598-749: its tbb loop over leaves with
// thread-local gradients summed at the end, :741-746): every rank evaluates the contribution of its leaves — uploading D2
// (and C) over its own PCIe link — and the 3 x natoms partial gradients are summed on the host in rank order.
int occ_xc_multi_gradient(occ_xc_multi *m, const double *D2, const double *C, double *gradient) {
  if (!m || !D2 || !gradient) return occ_xc_set_error("occ_xc_multi_gradient: null argument");
  const int n = m->n, natoms = m->h[0]->natoms;
  for (int r = 0; r < n; ++r) {
    if (!m->h[r]->gradient) return occ_xc_set_error("occ_xc_multi_gradient: handles were created without the gradient option");
    if (C && !m->h[r]->mo) return occ_xc_set_error("occ_xc_multi_gradient: the handles were created without occupied-orbital counts");
  }
  if (!m->h_grad &&
      cudaHostAlloc(&m->h_grad, 3 * static_cast<size_t>(natoms) * MAX_DEV * sizeof(double), cudaHostAllocPortable) != cudaSuccess)
    return occ_xc_set_error("occ_xc_multi_gradient: cudaHostAlloc failed");
  for (int r = 0; r < n; ++r)
    m->worker[r]->submit([=]() -> int {
      return xc_gradient_impl(m->h[r], D2, C, m->h_grad + 3 * static_cast<size_t>(natoms) * r);
    });
  std::string err;
  int rc = 0;
  for (int r = 0; r < n; ++r) rc |= m->worker[r]->wait(&err);
  if (rc) return occ_xc_set_error(err.c_str());
  for (int k = 0; k < 3 * natoms; ++k) {
    double g = 0.0;
    for (int r = 0; r < n; ++r) g += m->h_grad[3 * static_cast<size_t>(natoms) * r + k];
    gradient[k] = g;
  }
  return 0;
}

} // extern "C"
