// ctx.cuh -- the context behind the C-ABI (device buffers, streams, state).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "common.cuh"

namespace mcov {

// grow-only device / pinned buffers
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool pageable = false;            // plain malloc instead of pinned memory (pinning costs 0.6 - 2.4 ms per MB on these boxes)
  cudaError_t ensure(size_t bytes, bool pinned = true) {
    if (bytes <= cap && pageable == !pinned) return cudaSuccess;
    release();
    size_t want = bytes + bytes / 8 + 256;
    if (pinned) {
      cudaError_t e = cudaMallocHost(&p, want);
      if (e != cudaSuccess) { p = nullptr; return e; }
    } else {
      p = std::malloc(want);
      if (!p) return cudaErrorMemoryAllocation;
    }
    cap = want; pageable = !pinned;
    return cudaSuccess;
  }
  void release() { if (p) { if (pageable) std::free(p); else cudaFreeHost(p); } p = nullptr; cap = 0; pageable = false; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

enum PassState { kIdle = 0, kAccumulating = 1, kDepthReady = 2, kStreaming = 3 };

// kernel ids for the optional per-kernel CUDA-event timing (mcov_profile_*)
enum KernelId {
  kKExpand = 0, kKScan, kKFusedPrep, kKTileFirst, kKScanCounts, kKFarScatter, kKFusedTile,
  kKInitStats, kKRegionStats, kKWindowSums, kKIsizeHist, kKGroupCount, kKSortedStats, kKClear, kKRegionStatsSmall, kKCapReplay, kKUnpack, kKKmerHist, kKRegionStatsWarp,
  kKExpPrep, kKExpEntries, kKExpRegion, kKExpRevsum, kKRegionHist, kKHistFinish, kKRunCount, kKRunOffsets, kKRunWrite, kKRunEnds, kKBgzfInflate, kKBamGuess, kKBamWalkCount, kKBamWalkWrite, kKDeltaUnpack, kKStreamAcc, kKFusedPrepTma, kKFusedTileTma, kKStatsStream, kKStatsSplitFinish, kKBlockUnpack, kKBamNamesSeq,
  kKernelCount
};

struct ProfRec { int id; cudaEvent_t a, b; };

// cached chunk table of the last region set (mcov_region_stats_run)
struct RegionPlan {
  bool valid = false;
  int64_t g = 0, n_tasks = 0, n_small = 0;
  int32_t n_multi = 0;
  int32_t ss_grid = 0, n_split = 0;      // balanced stream over the large regions (k_stats_stream.cuh): CTAs, split regions
  int64_t n_contigs_epoch = -1;
  std::vector<int32_t> tid, start, end, rlen, rpad;
};

struct ReadStage {            // one staging set for host-resident batches
  DevBuf tid, pos, flag, mapq, cig_off, cig;
  DevBuf raw;                       // the narrow host transports land here before they are widened into the columns
  cudaEvent_t consumed = nullptr;   // recorded after the kernel that read this set
  bool in_flight = false;
};

}  // namespace mcov

struct mcov_ctx {
  int device = 0;
  int n_sm = 0;               // multiprocessors of the device (queried in mcov_create); grids are sized from it
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t d2h_stream = nullptr;   // copy-back of pipelined statistics records (overlaps the next pass)
  cudaEvent_t copied = nullptr;
  cudaStream_t unpack_stream = nullptr;   // the transport block's unpack kernel: depends on the block's copy only, so it may
  cudaEvent_t unpacked = nullptr;         // overlap the kernels of the previous pass on the main stream
  bool copy_pending = false;       // a transport block's copy has been enqueued and not yet waited for
  std::string err;

  int32_t n_contigs = 0;
  std::vector<int32_t> len;
  std::vector<int64_t> off;
  int64_t n_slots = 0;        // padded to a multiple of 4
  mcov::DevBuf d_len, d_off;

  int32_t* depth = nullptr;   // difference array, then depth
  mcov::DevBuf depth_own;
  bool depth_bound = false;

  mcov_filter filt;
  uint32_t flag_lut[128] = {0};   // the flag part of the filter as a 4096-bit table (rebuilt by mcov_set_filter)
  bool flag_lut_dirty = true;
  mcov::DevBuf d_flag_lut;        // its device copy
  int state = mcov::kIdle;

  mcov::DevBuf d_pc;          // PassCounters
  mcov::DevBuf d_status;      // scan tile status words
  mcov::ReadStage stage[2];
  int stage_next = 0;
  int64_t n_reads_pushed = 0;
  bool verdict_pending = false;   // an asynchronous fused pass has not been checked for sortedness yet
  int64_t contig_epoch = 0;
  mcov::RegionPlan plan;
  int32_t cap_contigs = 0;        // contigs replayed under htslib's max_depth cap in the last fused pass
  uint8_t* cap_flags = nullptr;   // device, [n_contigs]: contigs replayed in the last fused pass (inside d_status)
  std::vector<unsigned char> fused_blob;   // FusedArgs of the last fused pass (k_fused.cuh)

  // GPU-side BAM decode (bam_gpu.cu): compressed image, inflated stream, record-chain scratch, SoA columns
  struct BamDev {
    mcov::DevBuf raw, blocks, data, status, starts, segs, wout, tid, pos, flag, mapq, lseq, isize, cig_off, cig;
    mcov::DevBuf rec_off, name_hash, kmer, win;      // record offsets in `data`; read names / SEQ derived from them on request
    // streamed decode (mcov_bam_gpu_stream_depth): pinned file chunks, the bytes of the record a chunk ended in, the reads
    // carried into the next batch
    mcov::PinBuf pin[2];
    mcov::DevBuf tail, tail2, c_tid, c_pos, c_flag, c_mapq, c_off, c_cig;
    int64_t n_rec = -1;                              // records of the last decode (-1: none)
    void release() {
      n_rec = -1;
      mcov::DevBuf* b[] = {&raw, &blocks, &data, &status, &starts, &segs, &wout, &tid, &pos, &flag, &mapq, &lseq, &isize, &cig_off, &cig,
                           &rec_off, &name_hash, &kmer, &win, &tail, &tail2, &c_tid, &c_pos, &c_flag, &c_mapq, &c_off, &c_cig};
      for (mcov::DevBuf* x : b) x->release();
      pin[0].release(); pin[1].release();
    }
  } bam;

  // fused (sorted) path scratch
  mcov::DevBuf d_end_slot, d_start_slot, d_far_list, d_tile_cnt, d_tile_off, d_far_sorted;

  // stats scratch
  mcov::DevBuf d_ss_pieces, d_ss_cta, d_ss_split, d_ss_pool;
  // streamed passes (mcov_stream_begin / mcov_stream_push)
  mcov::DevBuf d_stream_acc;          // StreamAcc + the carried reads' counts of the current batch
  mcov::DevBuf d_cap_scratch;         // per-region replays of the max_depth cap (mcov_region_stats_run)
  const void* ncig_key_ptr = nullptr; // device-resident offset array whose total op count is cached (stage_reads)
  int64_t ncig_key_n = -1, ncig_val = -1;
  int64_t stream_tile_lo = 0;         // tiles below this one hold final depth
  int64_t stream_reads = 0;           // distinct reads pushed so far
  bool stream_started = false;        // (count_del = 0 streams: the difference array has been cleared)
  mcov::DevBuf d_tasks, d_rlen, d_rchunks, d_rhist, d_pool, d_done, d_out, d_win_slot, d_win_n, d_win_out, d_htasks, d_tile_heavy, d_run_tasks, d_run_counts, d_run_out;
  int64_t n_runs = -1;               // records held in d_run_out ([tid | start | end | depth] x n_runs), -1 = none
  mcov::PinBuf h_pin;
  // pipelined statistics (mcov_region_stats_submit / collect): two pinned slots
  struct StatSlot {
    mcov::PinBuf buf;              // [g records | PassCounters]
    mcov::DevBuf d_rec;            // the slot's records on the device (copied back on d2h_stream)
    cudaEvent_t ready = nullptr;   // records written (main stream)
    cudaEvent_t done = nullptr;    // records in `buf` (d2h_stream)
    int64_t g = -1;                // -1 = nothing submitted
    bool has_verdict = false;
    bool has_subregion = false;    // some region starts inside its contig (matters only when the max_depth cap fired)
    int32_t cap_contigs = 0;       // of the pass the slot's records describe (-1: arrives with the verdict)
    int64_t n_reads = 0;
    std::vector<int32_t> len0;     // regions of length 0 (their records are zeroed, like mcov_region_stats_run)
  } slot[2];

  // launch accounting + optional per-kernel event timing
  int64_t n_launches = 0;
  bool profiling = false;
  std::vector<mcov::ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  double prof_ms[mcov::kKernelCount] = {0};
  int64_t prof_n[mcov::kKernelCount] = {0};

  cudaEvent_t ev_get() {
    if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  void prof_begin(int id) {
    ++n_launches;
    if (!profiling) return;
    mcov::ProfRec r; r.id = id; r.a = ev_get(); r.b = ev_get();
    cudaEventRecord(r.a, stream);
    prof.push_back(r);
  }
  void prof_end() {
    if (!profiling || prof.empty()) return;
    cudaEventRecord(prof.back().b, stream);
  }
  // fold finished records into the per-kernel totals (stream must be synchronised)
  void prof_collect() {
    for (auto& r : prof) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { prof_ms[r.id] += ms; prof_n[r.id] += 1; }
      ev_pool.push_back(r.a); ev_pool.push_back(r.b);
    }
    prof.clear();
  }
};

// wrap one kernel launch (or memset) for accounting / timing
#define MCOV_LAUNCH(ctx, id, ...) do { (ctx)->prof_begin(id); __VA_ARGS__; (ctx)->prof_end(); } while (0)
