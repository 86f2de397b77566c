// pl2gpu.cu - C-ABI entry points (include/plink2_b200.h): context, staging, KING job driver.
#include <cstdarg>
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/plink2_b200.h"
#include "common.cuh"
#include "king_kernels.cuh"
#include "king_ts_kernel.cuh"
#include "king_pairs_kernel.cuh"
#include "umma_probe.cuh"

namespace pl2 {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// ---- tile list over the strict lower triangle restricted to rows [row_start,row_end) ----
static uint32_t ColTilesForRowTile(uint32_t rt, uint32_t row_end, bool include_diag, uint32_t tile_cols) {
  uint32_t tile_row_end = (rt + 1) * kTileRows;
  if (tile_row_end > row_end) tile_row_end = row_end;
  // columns 0 .. tile_row_end-2 are needed (.. tile_row_end-1 with the diagonal)
  const uint32_t cols = include_diag ? tile_row_end : (tile_row_end ? tile_row_end - 1 : 0);
  if (!cols) return 0;
  return DivUpU32(cols, tile_cols);
}

uint64_t CountTiles(uint32_t row_start, uint32_t row_end, bool include_diag, uint32_t tile_cols) {
  if (row_end <= row_start) return 0;
  uint64_t n = 0;
  for (uint32_t rt = row_start / kTileRows; rt * kTileRows < row_end; ++rt) n += ColTilesForRowTile(rt, row_end, include_diag, tile_cols);
  return n;
}

int BuildTileList(uint32_t row_start, uint32_t row_end, bool include_diag, TileList* tl, uint32_t tile_cols) {
  std::vector<uint32_t> rt_v, tc_v, off_v;
  tl->row_tile_first = row_start / kTileRows;
  uint32_t rt = tl->row_tile_first;
  for (; rt * kTileRows < row_end; ++rt) {
    off_v.push_back(static_cast<uint32_t>(rt_v.size()));
    const uint32_t nct = ColTilesForRowTile(rt, row_end, include_diag, tile_cols);
    for (uint32_t tc = 0; tc < nct; ++tc) {
      rt_v.push_back(rt);
      tc_v.push_back(tc);
    }
  }
  off_v.push_back(static_cast<uint32_t>(rt_v.size()));
  tl->row_tile_ct = rt - tl->row_tile_first;
  tl->tile_ct = static_cast<uint32_t>(rt_v.size());
  const size_t nb = (rt_v.size() + 1) * sizeof(uint32_t);
  PL2_CUDA_OK(cudaMalloc(&tl->d_tile_rt, nb));
  PL2_CUDA_OK(cudaMalloc(&tl->d_tile_tc, nb));
  PL2_CUDA_OK(cudaMalloc(&tl->d_rowtile_offset, off_v.size() * sizeof(uint32_t)));
  if (!rt_v.empty()) {
    PL2_CUDA_OK(cudaMemcpy(tl->d_tile_rt, rt_v.data(), rt_v.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    PL2_CUDA_OK(cudaMemcpy(tl->d_tile_tc, tc_v.data(), tc_v.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  PL2_CUDA_OK(cudaMemcpy(tl->d_rowtile_offset, off_v.data(), off_v.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  tl->h_rowtile_offset = off_v;
  // launch order: blocks of kBand x kBand tiles (~ one wave of 148 CTAs) so that the CTAs resident at
  // the same time stream the same few row/column sample ranges and hit in L2
  constexpr uint32_t kBand = 12;
  std::vector<uint32_t> order;
  order.reserve(rt_v.size());
  const uint32_t n_rt = tl->row_tile_ct;
  for (uint32_t rb = 0; rb < n_rt; rb += kBand) {
    const uint32_t rb_end = std::min(n_rt, rb + kBand);
    uint32_t max_cols = 0;
    for (uint32_t r = rb; r < rb_end; ++r) max_cols = std::max(max_cols, off_v[r + 1] - off_v[r]);
    for (uint32_t cb = 0; cb < max_cols; cb += kBand) {
      for (uint32_t r = rb; r < rb_end; ++r) {
        const uint32_t ncols = off_v[r + 1] - off_v[r];
        for (uint32_t c = cb; c < std::min(ncols, cb + kBand); ++c) order.push_back(off_v[r] + c);
      }
    }
  }
  PL2_CUDA_OK(cudaMalloc(&tl->d_tile_order, nb));
  if (!order.empty()) PL2_CUDA_OK(cudaMemcpy(tl->d_tile_order, order.data(), order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  return 0;
}

void FreeTileList(TileList* tl) {
  cudaFree(tl->d_tile_rt);
  cudaFree(tl->d_tile_tc);
  cudaFree(tl->d_rowtile_offset);
  cudaFree(tl->d_tile_order);
  tl->d_tile_order = nullptr;
  tl->d_tile_rt = tl->d_tile_tc = tl->d_rowtile_offset = nullptr;
}

// ---- staged genotype block on the device ----
int StageAlloc(uint32_t sample_ct, uint32_t variant_cap, GenoStage* gs, uint32_t sample_pad) {
  gs->sample_ct = sample_ct;
  gs->sample_ct_padded = RoundUpU32(sample_ct, sample_pad);
  gs->pitch = gs->sample_ct_padded / 4;
  gs->variant_cap = RoundUpU32(variant_cap, kVariantPad);
  if (cudaMalloc(&gs->d_raw, static_cast<uint64_t>(gs->variant_cap) * gs->pitch) != cudaSuccess) {
    cudaGetLastError();
    gs->d_raw = nullptr;
    set_error("insufficient device memory for a %u-variant x %u-sample genotype stage", gs->variant_cap, sample_ct);
    return 1;
  }
  return 0;
}

int LaunchPadGenotypes(Ctx* ctx, uint8_t* dst, uint32_t pitch, uint32_t sample_ct, uint32_t variant_ct, uint32_t variant_ct_padded) {
  if (!variant_ct_padded) return 0;
  pad_genotypes_kernel<<<variant_ct_padded, 128, 0, ctx->stream>>>(dst, pitch, sample_ct, variant_ct, variant_ct_padded);
  ctx->launches++;
  PL2_CUDA_OK(cudaGetLastError());
  return 0;
}

void StageFree(GenoStage* gs) {
  cudaFree(gs->d_raw);
  gs->d_raw = nullptr;
}

int StageUpload(Ctx* ctx, GenoStage* gs, const void* src, uint64_t src_stride, uint32_t variant_ct, int src_is_device, uint32_t* padded_ct_ptr, uint32_t dst_row, uint32_t pad_to) {
  const uint32_t padded = RoundUpU32(variant_ct, pad_to);
  if (dst_row + padded > gs->variant_cap) {
    set_error("StageUpload: %u + %u rows exceed the stage capacity %u", dst_row, padded, gs->variant_cap);
    return 1;
  }
  const uint32_t width = DivUpU32(gs->sample_ct, 4);
  uint8_t* dst = gs->d_raw + static_cast<uint64_t>(dst_row) * gs->pitch;
  if (variant_ct) {
    PL2_CUDA_OK(cudaMemcpy2DAsync(dst, gs->pitch, src, src_stride, width, variant_ct, src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  }
  if (padded) {
    pad_genotypes_kernel<<<padded, 128, 0, ctx->stream>>>(dst, gs->pitch, gs->sample_ct, variant_ct, padded);
    ctx->launches++;
    PL2_CUDA_OK(cudaGetLastError());
  }
  *padded_ct_ptr = padded;
  return 0;
}

}  // namespace pl2

using namespace pl2;

struct Pl2KingJob {
  Pl2GpuCtx* ctx = nullptr;
  uint32_t sample_ct = 0, row_start = 0, row_end = 0;
  int algo = kPl2KingAlgoTensor;
  TileList tiles;
  GenoStage stage;
  uint32_t* d_planes = nullptr;  // popcount path only
  uint8_t* d_raw_t = nullptr;    // TS path only: row-side re-tiled copy of the staged block (king_ts_kernel.cuh)
  uint8_t* d_raw_j = nullptr;    // TS path only: column-side re-tiled copy
  uint32_t tile_cols = kTileCols;
  int32_t* d_raw_acc = nullptr;
  void* d_out_stage = nullptr;   // bounded staging for host downloads
  uint64_t out_stage_bytes = 0;
  uint64_t variants_added = 0;
  // TS path, host sources: H2D of batch k+1 on the copy stream overlaps the tensor kernel of batch k
  cudaEvent_t ev_copied = nullptr;      // the staged block has arrived (copy stream)
  cudaEvent_t ev_stage_free = nullptr;  // the staged block has been re-tiled and may be overwritten (compute stream)
};

extern "C" {

int pl2gpu_abi_version(void) { return 1; }

const char* pl2gpu_last_error(void) { return get_error(); }

int pl2gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int pl2gpu_ctx_create(int device_idx, Pl2GpuCtx** ctx_ptr) {
  *ctx_ptr = nullptr;
  int n = 0;
  PL2_CUDA_OK(cudaGetDeviceCount(&n));
  if (device_idx < 0 || device_idx >= n) {
    set_error("pl2gpu_ctx_create: device %d out of range (%d CUDA devices visible); there is no CPU fallback", device_idx, n);
    return 1;
  }
  cudaDeviceProp prop;
  PL2_CUDA_OK(cudaGetDeviceProperties(&prop, device_idx));
  if (prop.major != 10) {
    set_error("pl2gpu_ctx_create: device %d is sm_%d%d; this library contains sm_100a code only", device_idx, prop.major, prop.minor);
    return 1;
  }
  PL2_CUDA_OK(cudaSetDevice(device_idx));
  Pl2GpuCtx* ctx = new Pl2GpuCtx();
  ctx->c.device = device_idx;
  ctx->c.sm_count = prop.multiProcessorCount;
  PL2_CUDA_OK(cudaStreamCreateWithFlags(&ctx->c.stream, cudaStreamNonBlocking));
  PL2_CUDA_OK(cudaStreamCreateWithFlags(&ctx->c.copy_stream, cudaStreamNonBlocking));
  PL2_CUDA_OK(cudaFuncSetAttribute(king_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  PL2_CUDA_OK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kProbeSmemBytes));
  PL2_CUDA_OK(cudaFuncSetAttribute(umma_probe_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kProbeSmemBytes));
  PL2_CUDA_OK(cudaFuncSetAttribute(king_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmemBytes));
  *ctx_ptr = ctx;
  return 0;
}

int pl2gpu_ctx_destroy(Pl2GpuCtx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->c.device);
  if (ctx->c.stream) cudaStreamDestroy(ctx->c.stream);
  if (ctx->c.copy_stream) cudaStreamDestroy(ctx->c.copy_stream);
  for (auto& e : ctx->c.events)
    if (e) cudaEventDestroy(e);
  delete ctx;
  return 0;
}

int pl2gpu_ctx_synchronize(Pl2GpuCtx* ctx) {
  PL2_CUDA_OK(cudaSetDevice(ctx->c.device));
  PL2_CUDA_OK(cudaStreamSynchronize(ctx->c.stream));
  return 0;
}

void* pl2gpu_ctx_stream(Pl2GpuCtx* ctx) { return ctx ? static_cast<void*>(ctx->c.stream) : nullptr; }

uint64_t pl2gpu_ctx_launch_count(Pl2GpuCtx* ctx) { return ctx ? ctx->c.launches : 0; }

int pl2gpu_ctx_mem_info(Pl2GpuCtx* ctx, uint64_t* free_bytes, uint64_t* total_bytes) {
  if (!ctx) {
    set_error("pl2gpu_ctx_mem_info: null context");
    return 1;
  }
  PL2_CUDA_OK(cudaSetDevice(ctx->c.device));
  size_t f = 0, t = 0;
  PL2_CUDA_OK(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return 0;
}

int pl2gpu_host_alloc(uint64_t bytes, void** ptr) {
  *ptr = nullptr;
  PL2_CUDA_OK(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}

int pl2gpu_host_free(void* ptr) {
  if (ptr) PL2_CUDA_OK(cudaFreeHost(ptr));
  return 0;
}

int pl2gpu_ctx_event_record(Pl2GpuCtx* ctx, int slot) {
  if (!ctx || slot < 0 || slot >= 16) {
    set_error("pl2gpu_ctx_event_record: bad arguments");
    return 1;
  }
  PL2_CUDA_OK(cudaSetDevice(ctx->c.device));
  if (!ctx->c.events[slot]) PL2_CUDA_OK(cudaEventCreate(&ctx->c.events[slot]));
  PL2_CUDA_OK(cudaEventRecord(ctx->c.events[slot], ctx->c.stream));
  return 0;
}

int pl2gpu_ctx_event_elapsed_ms(Pl2GpuCtx* ctx, int slot_from, int slot_to, float* ms) {
  if (!ctx || slot_from < 0 || slot_from >= 16 || slot_to < 0 || slot_to >= 16 || !ctx->c.events[slot_from] || !ctx->c.events[slot_to]) {
    set_error("pl2gpu_ctx_event_elapsed_ms: bad arguments");
    return 1;
  }
  PL2_CUDA_OK(cudaSetDevice(ctx->c.device));
  PL2_CUDA_OK(cudaEventSynchronize(ctx->c.events[slot_to]));
  PL2_CUDA_OK(cudaEventElapsedTime(ms, ctx->c.events[slot_from], ctx->c.events[slot_to]));
  return 0;
}

// ------------------------------------------------------------------------------------------ KING

uint64_t pl2gpu_king_mem_required(uint32_t sample_ct, uint32_t row_start, uint32_t row_end, uint32_t max_variants_per_add) {
  // Upper bound over the algorithms (the caller does not pass one): accumulators + staged block and
  // its per-algorithm re-layouts + tile lists + slack.
  uint32_t cap = max_variants_per_add ? max_variants_per_add : kMaxStageVariants;
  if (cap > kMaxStageVariants) cap = kMaxStageVariants;
  cap = RoundUpU32(cap, kVariantPad);
  const uint64_t slack = 256ull << 20;
  // SS tensor / popcount: 128 x 96 tiles, raw block + 3 bit planes (popcount only)
  const uint64_t tiles = CountTiles(row_start, row_end, false);
  const uint64_t npad = RoundUpU32(sample_ct, kSamplePad);
  const uint64_t need_ss = tiles * kKingTileAccWords * 4 + static_cast<uint64_t>(cap) * (npad / 4) + 3ull * (cap / 32) * npad * 4 + tiles * 12;
  // TS tensor (the default): 128 x 80 tiles, raw block + row-side and column-side re-tiled copies
  const uint64_t tiles_ts = CountTiles(row_start, row_end, false, kTsCols);
  const uint64_t npad_ts = RoundUpU32(sample_ct, kTsSamplePad);
  const uint64_t need_ts = tiles_ts * kTsTileAccWords * 4 + 3ull * cap * (npad_ts / 4) + tiles_ts * 12;
  return (need_ss > need_ts ? need_ss : need_ts) + slack;
}

int pl2gpu_king_begin(Pl2GpuCtx* ctx, uint32_t sample_ct, uint32_t row_start, uint32_t row_end, int algo, Pl2KingJob** job_ptr) {
  *job_ptr = nullptr;
  if (!ctx) {
    set_error("pl2gpu_king_begin: null context");
    return 1;
  }
  if (sample_ct < 2 || row_end > sample_ct || row_start >= row_end) {
    set_error("pl2gpu_king_begin: bad row range [%u,%u) for %u samples", row_start, row_end, sample_ct);
    return 1;
  }
  if (algo == kPl2KingAlgoAuto) algo = kPl2KingAlgoTensorTS;
  if (algo != kPl2KingAlgoPopcount && algo != kPl2KingAlgoTensor && algo != kPl2KingAlgoTensorTS) {
    set_error("pl2gpu_king_begin: unknown algo %d", algo);
    return 1;
  }
  PL2_CUDA_OK(cudaSetDevice(ctx->c.device));
  Pl2KingJob* job = new Pl2KingJob();
  job->ctx = ctx;
  job->sample_ct = sample_ct;
  job->row_start = row_start;
  job->row_end = row_end;
  job->algo = algo;
  auto fail = [&]() {
    pl2gpu_king_end(job);
    return 1;
  };
  const bool ts = algo == kPl2KingAlgoTensorTS;
  if (cudaEventCreateWithFlags(&job->ev_copied, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&job->ev_stage_free, cudaEventDisableTiming) != cudaSuccess) {
    set_error("pl2gpu_king_begin: cudaEventCreate failed");
    return fail();
  }
  job->tile_cols = ts ? kTsCols : kTileCols;
  if (BuildTileList(row_start, row_end, false, &job->tiles, job->tile_cols)) return fail();
  if (StageAlloc(sample_ct, kMaxStageVariants, &job->stage, ts ? kTsSamplePad : kSamplePad)) return fail();
  if (ts && (cudaMalloc(&job->d_raw_t, static_cast<uint64_t>(job->stage.sample_ct_padded) * (job->stage.variant_cap / 4)) != cudaSuccess ||
             cudaMalloc(&job->d_raw_j, static_cast<uint64_t>(job->stage.sample_ct_padded) * (job->stage.variant_cap / 4)) != cudaSuccess)) {
    cudaGetLastError();
    set_error("pl2gpu_king_begin: insufficient device memory for the re-tiled genotype copies");
    return fail();
  }
  const uint64_t acc_bytes = static_cast<uint64_t>(job->tiles.tile_ct) * (5ull * job->tile_cols * kTileRows) * sizeof(int32_t);
  if (cudaMalloc(&job->d_raw_acc, acc_bytes ? acc_bytes : 4) != cudaSuccess) {
    cudaGetLastError();
    set_error("pl2gpu_king_begin: insufficient device memory for %u pair tiles (%.1f GB of accumulators); narrow the row range", job->tiles.tile_ct, acc_bytes / 1e9);
    return fail();
  }
  if (cudaMemsetAsync(job->d_raw_acc, 0, acc_bytes, ctx->c.stream) != cudaSuccess) {
    set_error("pl2gpu_king_begin: cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail();
  }
  if (algo == kPl2KingAlgoPopcount) {
    const uint64_t plane_bytes = 3ull * (job->stage.variant_cap / 32) * job->stage.sample_ct_padded * sizeof(uint32_t);
    if (cudaMalloc(&job->d_planes, plane_bytes) != cudaSuccess) {
      cudaGetLastError();
      set_error("pl2gpu_king_begin: insufficient device memory for bit planes");
      return fail();
    }
  }
  job->out_stage_bytes = 256ull << 20;
  if (cudaMalloc(&job->d_out_stage, job->out_stage_bytes) != cudaSuccess) {
    cudaGetLastError();
    set_error("pl2gpu_king_begin: insufficient device memory for output staging");
    return fail();
  }
  *job_ptr = job;
  return 0;
}

int pl2gpu_king_add_variants(Pl2KingJob* job, const void* genovecs, uint64_t variant_stride_bytes, uint32_t variant_ct, int src_is_device) {
  if (!job) {
    set_error("pl2gpu_king_add_variants: null job");
    return 1;
  }
  Ctx* c = &job->ctx->c;
  PL2_CUDA_OK(cudaSetDevice(c->device));
  const uint64_t min_stride = 8ull * DivUpU32(job->sample_ct, 32);
  if (variant_stride_bytes < DivUpU32(job->sample_ct, 4)) {
    set_error("pl2gpu_king_add_variants: variant stride %llu < %u bytes of genotype data (PgrGet rows are %llu bytes)", static_cast<unsigned long long>(variant_stride_bytes), DivUpU32(job->sample_ct, 4), static_cast<unsigned long long>(min_stride));
    return 1;
  }
  const uint8_t* src = static_cast<const uint8_t*>(genovecs);
  uint32_t done = 0;
  while (done < variant_ct) {
    uint32_t cur = variant_ct - done;
    if (cur > job->stage.variant_cap) cur = job->stage.variant_cap;
    uint32_t padded = 0;
    // The stage buffer is reused: make sure the previous chunk's kernels are ordered before the
    // copy (same stream => implicit).
    const bool overlap_h2d = !src_is_device && job->algo == kPl2KingAlgoTensorTS && job->tiles.tile_ct;
    if (overlap_h2d) {
      // The TS kernel reads only the re-tiled copies, so the raw stage is free as soon as the previous
      // batch's re-tiling kernels are done: copy on the copy stream while the previous tensor kernel runs.
      padded = RoundUpU32(cur, kVariantPad);
      const uint32_t width = DivUpU32(job->stage.sample_ct, 4);
      PL2_CUDA_OK(cudaStreamWaitEvent(c->copy_stream, job->ev_stage_free, 0));
      PL2_CUDA_OK(cudaMemcpy2DAsync(job->stage.d_raw, job->stage.pitch, src + static_cast<uint64_t>(done) * variant_stride_bytes, variant_stride_bytes, width, cur, cudaMemcpyHostToDevice, c->copy_stream));
      PL2_CUDA_OK(cudaEventRecord(job->ev_copied, c->copy_stream));
      PL2_CUDA_OK(cudaStreamWaitEvent(c->stream, job->ev_copied, 0));
      PL2_TRY(LaunchPadGenotypes(c, job->stage.d_raw, job->stage.pitch, job->stage.sample_ct, cur, padded));
    } else {
      PL2_TRY(StageUpload(c, &job->stage, src + static_cast<uint64_t>(done) * variant_stride_bytes, variant_stride_bytes, cur, src_is_device, &padded));
    }
    if (job->tiles.tile_ct) {
      if (job->algo == kPl2KingAlgoPopcount) {
        const uint32_t word_ct = padded / 32;
        const uint64_t warps = static_cast<uint64_t>(job->stage.sample_ct_padded / 32) * word_ct;
        split_transpose_kernel<<<static_cast<uint32_t>(DivUpU64(warps, 8)), 256, 0, c->stream>>>(job->stage.d_raw, job->stage.pitch, job->stage.sample_ct_padded, word_ct, job->d_planes);
        c->launches++;
        king_popc_kernel<<<job->tiles.tile_ct * 2, 256, 0, c->stream>>>(job->d_planes, job->stage.sample_ct_padded, word_ct, job->tiles.d_tile_rt, job->tiles.d_tile_tc, job->d_raw_acc);
        c->launches++;
      } else if (job->algo == kPl2KingAlgoTensorTS) {
        const uint32_t coltile_ct = job->stage.sample_ct_padded / kTsCols;
        geno_tile_rows_kernel<<<dim3(padded / 64, job->stage.sample_ct_padded / 64), 256, 0, c->stream>>>(job->stage.d_raw, job->stage.pitch, padded / 32, job->d_raw_t);
        c->launches++;
        geno_tile_cols_kernel<<<dim3(padded / kTsKcJ, DivUpU32(coltile_ct, 16)), 256, 0, c->stream>>>(job->stage.d_raw, job->stage.pitch, padded / kTsKcJ, coltile_ct, job->d_raw_j);
        c->launches++;
        PL2_CUDA_OK(cudaEventRecord(job->ev_stage_free, c->stream));
        king_ts_kernel<<<job->tiles.tile_ct, kTsThreads, kTsSmemBytes, c->stream>>>(job->d_raw_j, job->d_raw_t, padded, job->tiles.d_tile_order, job->tiles.d_tile_rt, job->tiles.d_tile_tc, job->d_raw_acc);
        c->launches++;
      } else {
        king_tc_kernel<<<job->tiles.tile_ct, kTcThreads, kTcSmemBytes, c->stream>>>(job->stage.d_raw, job->stage.pitch, padded, job->tiles.d_tile_order, job->tiles.d_tile_rt, job->tiles.d_tile_tc, job->d_raw_acc);
        c->launches++;
      }
      PL2_CUDA_OK(cudaGetLastError());
    }
    if (overlap_h2d) {
      PL2_CUDA_OK(cudaEventSynchronize(job->ev_copied));  // the caller may reuse its buffer; the kernels keep running
    } else if (!src_is_device) {
      // host source on the compute stream: it has been consumed once the stream reaches here
      PL2_CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    done += cur;
  }
  job->variants_added += variant_ct;
  return 0;
}

static int KingGet(Pl2KingJob* job, uint32_t r0, uint32_t r1, void* dst, int dst_is_device, bool kinship) {
  if (!job) {
    set_error("pl2gpu_king_get: null job");
    return 1;
  }
  if (r0 < job->row_start || r1 > job->row_end || r0 > r1) {
    set_error("pl2gpu_king_get: rows [%u,%u) outside the job's [%u,%u)", r0, r1, job->row_start, job->row_end);
    return 1;
  }
  Ctx* c = &job->ctx->c;
  PL2_CUDA_OK(cudaSetDevice(c->device));
  const uint64_t bytes_per_pair = kinship ? 8 : 20;
  auto tri = [](uint64_t r) { return r ? r * (r - 1) / 2 : 0ull; };
  uint8_t* out = static_cast<uint8_t*>(dst);
  uint32_t cur0 = r0;
  while (cur0 < r1) {
    uint32_t cur1;
    void* d_dst;
    if (dst_is_device) {
      cur1 = r1;
      d_dst = out;
    } else {
      // largest row block whose pairs fit the staging buffer (at least one row)
      cur1 = cur0 + 1;
      while (cur1 < r1 && (tri(cur1 + 1) - tri(cur0)) * bytes_per_pair <= job->out_stage_bytes) ++cur1;
      if ((tri(cur1) - tri(cur0)) * bytes_per_pair > job->out_stage_bytes) {
        set_error("pl2gpu_king_get: a single row exceeds the staging buffer");
        return 1;
      }
      d_dst = job->d_out_stage;
    }
    const uint64_t pairs = tri(cur1) - tri(cur0);
    if (pairs) {
      const uint32_t rt_a = cur0 / kTileRows - job->tiles.row_tile_first;
      const uint32_t rt_b = (cur1 - 1) / kTileRows - job->tiles.row_tile_first;
      const uint32_t tile_a = job->tiles.h_rowtile_offset[rt_a];
      const uint32_t tile_b = job->tiles.h_rowtile_offset[rt_b + 1];
      if (tile_b > tile_a) {
        const uint32_t grid = (tile_b - tile_a) * 8;
        const int32_t* acc0 = job->d_raw_acc + static_cast<uint64_t>(tile_a) * (5ull * job->tile_cols * kTileRows);
        const uint32_t* trt = job->tiles.d_tile_rt + tile_a;
        const uint32_t* ttc = job->tiles.d_tile_tc + tile_a;
        if (job->tile_cols == kTsCols) {
          if (kinship) king_finalize_kernel<true, kTsCols><<<grid, 256, 0, c->stream>>>(acc0, trt, ttc, job->sample_ct, cur0, cur1, nullptr, static_cast<double*>(d_dst));
          else king_finalize_kernel<false, kTsCols><<<grid, 256, 0, c->stream>>>(acc0, trt, ttc, job->sample_ct, cur0, cur1, static_cast<uint32_t*>(d_dst), nullptr);
        } else {
          if (kinship) king_finalize_kernel<true, kTileCols><<<grid, 256, 0, c->stream>>>(acc0, trt, ttc, job->sample_ct, cur0, cur1, nullptr, static_cast<double*>(d_dst));
          else king_finalize_kernel<false, kTileCols><<<grid, 256, 0, c->stream>>>(acc0, trt, ttc, job->sample_ct, cur0, cur1, static_cast<uint32_t*>(d_dst), nullptr);
        }
        c->launches++;
        PL2_CUDA_OK(cudaGetLastError());
      }
      if (!dst_is_device) {
        PL2_CUDA_OK(cudaMemcpyAsync(out, d_dst, pairs * bytes_per_pair, cudaMemcpyDeviceToHost, c->stream));
        PL2_CUDA_OK(cudaStreamSynchronize(c->stream));
        out += pairs * bytes_per_pair;
      }
    }
    cur0 = cur1;
  }
  if (dst_is_device) {
    // caller synchronises through pl2gpu_ctx_synchronize / its own stream ordering
  }
  return 0;
}

int pl2gpu_king_get_counts(Pl2KingJob* job, uint32_t out_row_start, uint32_t out_row_end, uint32_t* dst, int dst_is_device) {
  return KingGet(job, out_row_start, out_row_end, dst, dst_is_device, false);
}

int pl2gpu_king_get_kinship(Pl2KingJob* job, uint32_t out_row_start, uint32_t out_row_end, double* dst, int dst_is_device) {
  return KingGet(job, out_row_start, out_row_end, dst, dst_is_device, true);
}

int pl2gpu_king_get_filtered(Pl2KingJob* job, uint32_t r0, uint32_t r1, double min_kinship, uint64_t max_out, uint32_t* pairs_out, uint32_t* counts_out, double* kinship_out, uint64_t* n_found) {
  if (!job || !n_found || (max_out && (!pairs_out || !counts_out || !kinship_out))) {
    set_error("pl2gpu_king_get_filtered: bad arguments");
    return 1;
  }
  if (r0 < job->row_start || r1 > job->row_end || r0 > r1) {
    set_error("pl2gpu_king_get_filtered: rows [%u,%u) outside the job's [%u,%u)", r0, r1, job->row_start, job->row_end);
    return 1;
  }
  *n_found = 0;
  Ctx* c = &job->ctx->c;
  PL2_CUDA_OK(cudaSetDevice(c->device));
  if (!job->tiles.tile_ct || r0 == r1) return 0;
  unsigned long long* d_found = nullptr;
  uint32_t *d_pairs = nullptr, *d_counts = nullptr;
  double* d_kin = nullptr;
  const uint64_t cap = max_out ? max_out : 1;
  int rc = 1;
  do {
    if (cudaMalloc(&d_found, 8) != cudaSuccess || cudaMalloc(&d_pairs, cap * 8) != cudaSuccess || cudaMalloc(&d_counts, cap * 20) != cudaSuccess || cudaMalloc(&d_kin, cap * 8) != cudaSuccess) {
      cudaGetLastError();
      set_error("pl2gpu_king_get_filtered: insufficient device memory for %llu result slots", static_cast<unsigned long long>(max_out));
      break;
    }
    if (cudaMemsetAsync(d_found, 0, 8, c->stream) != cudaSuccess) break;
    if (job->tile_cols == kTsCols) {
      king_filter_kernel<kTsCols><<<job->tiles.tile_ct, kTileRows, 0, c->stream>>>(job->d_raw_acc, job->tiles.d_tile_rt, job->tiles.d_tile_tc, job->sample_ct, r0, r1, min_kinship, max_out, d_found, d_pairs, d_counts, d_kin);
    } else {
      king_filter_kernel<kTileCols><<<job->tiles.tile_ct, kTileRows, 0, c->stream>>>(job->d_raw_acc, job->tiles.d_tile_rt, job->tiles.d_tile_tc, job->sample_ct, r0, r1, min_kinship, max_out, d_found, d_pairs, d_counts, d_kin);
    }
    c->launches++;
    unsigned long long found = 0;
    if (cudaGetLastError() != cudaSuccess || cudaMemcpyAsync(&found, d_found, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) {
      set_error("pl2gpu_king_get_filtered: %s", cudaGetErrorString(cudaGetLastError()));
      break;
    }
    *n_found = found;
    const uint64_t k = found < max_out ? found : max_out;
    if (k) {
      std::vector<uint32_t> hp(2 * k), hc(5 * k);
      std::vector<double> hk(k);
      if (cudaMemcpy(hp.data(), d_pairs, k * 8, cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemcpy(hc.data(), d_counts, k * 20, cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemcpy(hk.data(), d_kin, k * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("pl2gpu_king_get_filtered: %s", cudaGetErrorString(cudaGetLastError()));
        break;
      }
      std::vector<uint64_t> order(k);
      for (uint64_t q = 0; q < k; ++q) order[q] = q;
      std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return hp[2 * a] != hp[2 * b] ? hp[2 * a] < hp[2 * b] : hp[2 * a + 1] < hp[2 * b + 1]; });
      for (uint64_t q = 0; q < k; ++q) {
        const uint64_t src = order[q];
        pairs_out[2 * q] = hp[2 * src];
        pairs_out[2 * q + 1] = hp[2 * src + 1];
        memcpy(counts_out + 5 * q, &hc[5 * src], 20);
        kinship_out[q] = hk[src];
      }
    }
    rc = 0;
  } while (0);
  cudaFree(d_found);
  cudaFree(d_pairs);
  cudaFree(d_counts);
  cudaFree(d_kin);
  return rc;
}

uint64_t pl2gpu_king_variants_added(Pl2KingJob* job) { return job ? job->variants_added : 0; }

int pl2gpu_king_end(Pl2KingJob* job) {
  if (!job) return 0;
  if (job->ctx) {
    cudaSetDevice(job->ctx->c.device);
    cudaStreamSynchronize(job->ctx->c.stream);
  }
  if (job->ctx && job->ctx->c.copy_stream) cudaStreamSynchronize(job->ctx->c.copy_stream);
  if (job->ev_copied) cudaEventDestroy(job->ev_copied);
  if (job->ev_stage_free) cudaEventDestroy(job->ev_stage_free);
  FreeTileList(&job->tiles);
  StageFree(&job->stage);
  cudaFree(job->d_planes);
  cudaFree(job->d_raw_t);
  cudaFree(job->d_raw_j);
  cudaFree(job->d_raw_acc);
  cudaFree(job->d_out_stage);
  cudaGetLastError();
  delete job;
  return 0;
}

// ------------------------------------------------------------------------------------------ pair list

struct Pl2KingPairJob {
  Pl2GpuCtx* ctx = nullptr;
  uint32_t sample_ct = 0;
  uint64_t pair_ct = 0;
  GenoStage stage;
  uint8_t* d_raw_t = nullptr;   // sample-major 2-bit copy of the staged block
  uint32_t* d_pairs = nullptr;  // [pair][2]
  uint32_t* d_counts = nullptr; // [pair][5]
};

int pl2gpu_king_pairs_begin(Pl2GpuCtx* ctx, uint32_t sample_ct, const uint32_t* pairs_host, uint64_t pair_ct, Pl2KingPairJob** job_ptr) {
  if (job_ptr) *job_ptr = nullptr;
  if (!ctx || !job_ptr || !sample_ct || (pair_ct && !pairs_host)) {
    set_error("pl2gpu_king_pairs_begin: bad arguments");
    return 1;
  }
  for (uint64_t p = 0; p < 2 * pair_ct; ++p) {
    if (pairs_host[p] >= sample_ct) {
      set_error("pl2gpu_king_pairs_begin: pair %llu names sample %u of %u", static_cast<unsigned long long>(p / 2), pairs_host[p], sample_ct);
      return 1;
    }
  }
  PL2_CUDA_OK(cudaSetDevice(ctx->c.device));
  Pl2KingPairJob* job = new Pl2KingPairJob();
  job->ctx = ctx;
  job->sample_ct = sample_ct;
  job->pair_ct = pair_ct;
  auto fail = [&]() {
    pl2gpu_king_pairs_end(job);
    return 1;
  };
  if (StageAlloc(sample_ct, kMaxStageVariants, &job->stage, 64)) return fail();
  const uint64_t n_alloc = pair_ct ? pair_ct : 1;
  if (cudaMalloc(&job->d_raw_t, static_cast<uint64_t>(job->stage.sample_ct_padded) * (job->stage.variant_cap / 4)) != cudaSuccess || cudaMalloc(&job->d_pairs, n_alloc * 8) != cudaSuccess ||
      cudaMalloc(&job->d_counts, n_alloc * 20) != cudaSuccess) {
    cudaGetLastError();
    set_error("pl2gpu_king_pairs_begin: insufficient device memory for %llu pairs", static_cast<unsigned long long>(pair_ct));
    return fail();
  }
  if (cudaMemcpyAsync(job->d_pairs, pairs_host, pair_ct * 8, cudaMemcpyHostToDevice, ctx->c.stream) != cudaSuccess || cudaMemsetAsync(job->d_counts, 0, n_alloc * 20, ctx->c.stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->c.stream) != cudaSuccess) {
    set_error("pl2gpu_king_pairs_begin: %s", cudaGetErrorString(cudaGetLastError()));
    return fail();
  }
  *job_ptr = job;
  return 0;
}

int pl2gpu_king_pairs_add_variants(Pl2KingPairJob* job, const void* genovecs, uint64_t variant_stride_bytes, uint32_t variant_ct, int src_is_device) {
  if (!job || (!genovecs && variant_ct)) {
    set_error("pl2gpu_king_pairs_add_variants: bad arguments");
    return 1;
  }
  Ctx* c = &job->ctx->c;
  PL2_CUDA_OK(cudaSetDevice(c->device));
  const uint64_t min_stride = static_cast<uint64_t>(DivUpU32(job->sample_ct, 4));
  if (variant_stride_bytes < min_stride) {
    set_error("pl2gpu_king_pairs_add_variants: variant stride %llu < %llu bytes of genotype data", static_cast<unsigned long long>(variant_stride_bytes), static_cast<unsigned long long>(min_stride));
    return 1;
  }
  const uint8_t* src = static_cast<const uint8_t*>(genovecs);
  uint32_t done = 0;
  while (done < variant_ct) {
    uint32_t cur = variant_ct - done;
    if (cur > job->stage.variant_cap) cur = job->stage.variant_cap;
    uint32_t padded = 0;
    PL2_TRY(StageUpload(c, &job->stage, src + static_cast<uint64_t>(done) * variant_stride_bytes, variant_stride_bytes, cur, src_is_device, &padded));
    if (job->pair_ct) {
      const uint32_t pitch_t = padded / 4;
      geno_transpose_kernel<<<dim3(padded / 64, job->stage.sample_ct_padded / 64), 256, 0, c->stream>>>(job->stage.d_raw, job->stage.pitch, job->d_raw_t, pitch_t);
      king_pairs_kernel<<<static_cast<uint32_t>(DivUpU64(job->pair_ct, 8)), 256, 0, c->stream>>>(job->d_raw_t, pitch_t, padded / 32, job->d_pairs, job->pair_ct, job->d_counts);
      c->launches += 2;
      PL2_CUDA_OK(cudaGetLastError());
    }
    if (!src_is_device) PL2_CUDA_OK(cudaStreamSynchronize(c->stream));
    done += cur;
  }
  return 0;
}

int pl2gpu_king_pairs_get_counts(Pl2KingPairJob* job, uint64_t pair_start, uint64_t pair_end, uint32_t* dst, int dst_is_device) {
  if (!job || pair_start > pair_end || pair_end > job->pair_ct || (!dst && pair_end > pair_start)) {
    set_error("pl2gpu_king_pairs_get_counts: bad arguments");
    return 1;
  }
  Ctx* c = &job->ctx->c;
  PL2_CUDA_OK(cudaSetDevice(c->device));
  if (pair_end > pair_start) {
    PL2_CUDA_OK(cudaMemcpyAsync(dst, job->d_counts + 5 * pair_start, (pair_end - pair_start) * 20, dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  }
  PL2_CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

int pl2gpu_king_pairs_end(Pl2KingPairJob* job) {
  if (!job) return 0;
  if (job->ctx) {
    cudaSetDevice(job->ctx->c.device);
    cudaStreamSynchronize(job->ctx->c.stream);
  }
  StageFree(&job->stage);
  cudaFree(job->d_raw_t);
  cudaFree(job->d_pairs);
  cudaFree(job->d_counts);
  cudaGetLastError();
  delete job;
  return 0;
}

// ------------------------------------------------------------------------------------------ probe

int pl2gpu_debug_umma(Pl2GpuCtx* ctx, const uint8_t* a_img, uint32_t a_bytes, const uint8_t* b_img, uint32_t b_bytes, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t a_step_bytes, uint32_t b_step_bytes, uint32_t k_steps, uint32_t idesc, uint32_t n, int32_t* d_out_host) {
  if (!ctx) {
    set_error("pl2gpu_debug_umma: null context");
    return 1;
  }
  Ctx* c = &ctx->c;
  PL2_CUDA_OK(cudaSetDevice(c->device));
  UmmaProbeParams prm;
  prm.a_bytes = a_bytes;
  prm.b_bytes = b_bytes;
  prm.b_smem_off = RoundUpU32(a_bytes, 1024);
  if (prm.b_smem_off + b_bytes + 1024 > kProbeSmemBytes || n > 256 || (n & 15)) {
    set_error("pl2gpu_debug_umma: images too large or bad n");
    return 1;
  }
  prm.a_lbo = a_lbo;
  prm.a_sbo = a_sbo;
  prm.b_lbo = b_lbo;
  prm.b_sbo = b_sbo;
  prm.a_step_bytes = a_step_bytes;
  prm.b_step_bytes = b_step_bytes;
  prm.k_steps = k_steps;
  prm.idesc = idesc;
  prm.n = n;
  uint8_t *d_a = nullptr, *d_b = nullptr;
  int32_t* d_d = nullptr;
  PL2_CUDA_OK(cudaMalloc(&d_a, a_bytes));
  PL2_CUDA_OK(cudaMalloc(&d_b, b_bytes));
  PL2_CUDA_OK(cudaMalloc(&d_d, 128ull * n * 4));
  PL2_CUDA_OK(cudaMemcpyAsync(d_a, a_img, a_bytes, cudaMemcpyHostToDevice, c->stream));
  PL2_CUDA_OK(cudaMemcpyAsync(d_b, b_img, b_bytes, cudaMemcpyHostToDevice, c->stream));
  umma_probe_kernel<<<1, 128, kProbeSmemBytes, c->stream>>>(d_a, d_b, prm, d_d);
  c->launches++;
  PL2_CUDA_OK(cudaGetLastError());
  PL2_CUDA_OK(cudaMemcpyAsync(d_out_host, d_d, 128ull * n * 4, cudaMemcpyDeviceToHost, c->stream));
  PL2_CUDA_OK(cudaStreamSynchronize(c->stream));
  cudaFree(d_a);
  cudaFree(d_b);
  cudaFree(d_d);
  return 0;
}

int pl2gpu_selftest_umma(Pl2GpuCtx* ctx, int verbose) {
  if (getenv("PL2_UMMA_BENCH")) {  // diagnostic: tcgen05.mma issue / execution rate (umma_probe.cuh)
    Ctx* c = &ctx->c;
    long long* d_out = nullptr;
    PL2_CUDA_OK(cudaMalloc(&d_out, 64));
    PL2_CUDA_OK(cudaFuncSetAttribute(umma_issue_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kProbeSmemBytes));
    const uint32_t reps = 512;
    for (uint32_t mode = 0; mode < 4; ++mode)
      for (uint32_t issuers = 1; issuers <= 2; ++issuers)
        for (uint32_t n : {80u, 160u, 224u})
          for (uint32_t ce : {0u, 1u}) {
            if (n * issuers > 448) continue;
            umma_issue_bench_kernel<<<1, 128, kProbeSmemBytes, c->stream>>>(n, reps, mode, issuers, ce, d_out);
            long long h[8] = {0};
            PL2_CUDA_OK(cudaMemcpyAsync(h, d_out, 64, cudaMemcpyDeviceToHost, c->stream));
            PL2_CUDA_OK(cudaStreamSynchronize(c->stream));
            printf("umma_bench style=%s mode=%s issuers=%u n=%3u commit_every=%u: issue %.1f clk/mma, complete %.1f clk/mma (floor %.0f)\n", (mode >> 1) ? "elect" : "lane0", (mode & 1) ? "TS" : "SS", issuers, n, ce, double(h[0]) / reps, double(h[1]) / reps, n / 2.0 * issuers);
          }
    cudaFree(d_out);
  }
  // Production operand layout (geno_expand.cuh operand_offset), M=128, N=96, K=64 (two k-steps).
  const uint32_t M = 128, N = 96, K = 64;
  const uint32_t lbo_a = operand_lbo(M), lbo_b = operand_lbo(N);
  std::vector<uint8_t> a(M * K), b(N * K);
  std::vector<int8_t> av(M * K), bv(N * K);
  uint32_t seed = 12345;
  auto rnd = [&]() {
    seed = seed * 1664525u + 1013904223u;
    return static_cast<int8_t>((seed >> 24) % 7) - 3;
  };
  for (uint32_t m = 0; m < M; ++m)
    for (uint32_t k = 0; k < K; ++k) {
      const int8_t v = rnd();
      av[m * K + k] = v;
      a[operand_offset(k, m / 16, lbo_a) + (m % 16)] = static_cast<uint8_t>(v);
    }
  for (uint32_t n = 0; n < N; ++n)
    for (uint32_t k = 0; k < K; ++k) {
      const int8_t v = rnd();
      bv[n * K + k] = v;
      b[operand_offset(k, n / 16, lbo_b) + (n % 16)] = static_cast<uint8_t>(v);
    }
  std::vector<int32_t> d(128 * N);
  const uint32_t idesc = make_idesc_i8(M, N, true, true);
  PL2_TRY(pl2gpu_debug_umma(ctx, a.data(), M * K, b.data(), N * K, lbo_a, kCoreBytes, lbo_b, kCoreBytes, 4 * lbo_a, 4 * lbo_b, K / 32, idesc, N, d.data()));
  uint32_t bad = 0;
  for (uint32_t m = 0; m < M; ++m)
    for (uint32_t n = 0; n < N; ++n) {
      int32_t ref = 0;
      for (uint32_t k = 0; k < K; ++k) ref += static_cast<int32_t>(av[m * K + k]) * bv[n * K + k];
      if (ref != d[m * N + n]) {
        if (verbose && bad < 8) fprintf(stderr, "selftest_umma mismatch m=%u n=%u got=%d want=%d\n", m, n, d[m * N + n], ref);
        ++bad;
      }
    }
  if (bad) {
    set_error("pl2gpu_selftest_umma: %u of %u accumulator entries differ from the scalar reference", bad, M * N);
    return 1;
  }
  // TS form: same B image, A rows written to tensor memory with tcgen05.st (K-major, 4 K-bytes per column)
  {
    Ctx* c = &ctx->c;
    std::vector<uint8_t> a_rows(M * K);
    for (uint32_t m = 0; m < M; ++m)
      for (uint32_t k = 0; k < K; ++k) a_rows[m * K + k] = static_cast<uint8_t>(av[m * K + k]);
    UmmaProbeParams prm{};
    prm.b_bytes = N * K;
    prm.b_lbo = lbo_b;
    prm.b_sbo = kCoreBytes;
    prm.b_step_bytes = 4 * lbo_b;
    prm.k_steps = K / 32;
    prm.idesc = make_idesc_i8(M, N, false, true);
    prm.n = N;
    uint8_t *d_a = nullptr, *d_b = nullptr;
    int32_t* d_d = nullptr;
    PL2_CUDA_OK(cudaMalloc(&d_a, M * K));
    PL2_CUDA_OK(cudaMalloc(&d_b, N * K));
    PL2_CUDA_OK(cudaMalloc(&d_d, 128ull * N * 4));
    PL2_CUDA_OK(cudaMemcpyAsync(d_a, a_rows.data(), M * K, cudaMemcpyHostToDevice, c->stream));
    PL2_CUDA_OK(cudaMemcpyAsync(d_b, b.data(), N * K, cudaMemcpyHostToDevice, c->stream));
    umma_probe_ts_kernel<<<1, 128, kProbeSmemBytes, c->stream>>>(d_a, d_b, prm, d_d);
    c->launches++;
    PL2_CUDA_OK(cudaGetLastError());
    PL2_CUDA_OK(cudaMemcpyAsync(d.data(), d_d, 128ull * N * 4, cudaMemcpyDeviceToHost, c->stream));
    PL2_CUDA_OK(cudaStreamSynchronize(c->stream));
    cudaFree(d_a);
    cudaFree(d_b);
    cudaFree(d_d);
    bad = 0;
    for (uint32_t m = 0; m < M; ++m)
      for (uint32_t n = 0; n < N; ++n) {
        int32_t ref = 0;
        for (uint32_t k = 0; k < K; ++k) ref += static_cast<int32_t>(av[m * K + k]) * bv[n * K + k];
        if (ref != d[m * N + n]) {
          if (verbose && bad < 8) fprintf(stderr, "selftest_umma(TS) mismatch m=%u n=%u got=%d want=%d\n", m, n, d[m * N + n], ref);
          ++bad;
        }
      }
    if (bad) {
      set_error("pl2gpu_selftest_umma: TS form: %u of %u accumulator entries differ from the scalar reference", bad, M * N);
      return 3;
    }
  }
  return 0;
}

}  // extern "C"
