// Communicator behind the multi-GPU entry points (include/mpros_gpu.h: mpros_comm_*, mpros_map_set_occupancy_banded,
// mpros_plan_batch_sharded). One rank per GPU; two transports:
//   * NCCL (the product path on an 8 x B200 box): libnccl.so.2 is resolved at run time, so the single-GPU library
//     has no link-time dependency on it; every exchange is enqueued on the context's stream (no host synchronisation).
//   * "local": several ranks inside ONE process (one host thread and one context per rank, any devices, usually the
//     same one) exchanging through cudaMemcpyAsync behind a host barrier. It exists so that the banded algorithms
//     (seam union of the component labelling, carry rows of the distance transforms) run under the single-GPU test suite.
#pragma once
#include "mpros_internal.cuh"

struct mpros_comm
{
    int world = 1, rank = 0;
    int kind = 0;           // 0 = NCCL, 1 = local
    void *nccl = nullptr;   // ncclComm_t
    bool own_nccl = false;  // created by mpros_comm_create_nccl (destroyed with the handle) vs adopted from the caller
    struct LocalGroup *grp = nullptr;
    uint64_t bytes_sent = 0, exchanges = 0; // payload this rank contributed / number of exchange steps (reported by tests and bench)
};

namespace mpros
{
// In-place all-gather with per-rank extents: rank r's part lives at buf + off[r] (cnt[r] bytes) on every rank; after the
// call every rank holds every part. Stream-ordered on c->stream for NCCL; the local transport synchronises the host threads.
int comm_allgatherv(Ctx *c, mpros_comm *cm, void *buf, const size_t *off, const size_t *cnt);
// Equal-size form: part r at buf + r * bytes.
int comm_allgather(Ctx *c, mpros_comm *cm, void *buf, size_t bytes);
// Rows [lo_r, hi_r) of a row-major layer with `row_bytes` per row; lo/hi per rank from band_of().
int comm_allgather_rows(Ctx *c, mpros_comm *cm, void *layer, size_t row_bytes, int total_rows, int rows_per_rank);
inline void band_of(int total_rows, int rows_per_rank, int rank, int *lo, int *hi)
{
    long long a = (long long)rank * rows_per_rank;
    *lo = (int)(a < total_rows ? a : total_rows);
    long long b = (long long)*lo + rows_per_rank;
    *hi = (int)(b < total_rows ? b : total_rows);
}
int map_build_voronoi_banded(Ctx *c, mpros_comm *cm, int rows_per_rank); // N5-N9 on row bands, voronoi.cu
// one planner launch drawing tickets from `counter` and storing into the given result arrays, any of which may be peer memory (plan.cu)
int plan_batch_shared_launch(mpros_ctx *c, const double *d_starts, const double *d_goals, int nq, unsigned int *counter, int32_t *status,
                             double *cost, int32_t *path_len, double *paths, int path_cap, int64_t *stats);
} // namespace mpros
