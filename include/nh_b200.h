/*
 * nh_b200.h -- C ABI of libnh_b200.so, the sm_100a kernel library behind
 * nano-hevc's block-coding hot path.
 *
 * The reference (Luodian/nano-hevc) is pure Python + numpy and has no FFI of
 * its own: its "operator API" is the flat set of module-level functions
 * re-exported by nano_hevc/__init__.py:5-48.  Every entry point below names the
 * reference function (file:line, relative to the reference checkout) it
 * replaces; INTEGRATION.md shows the ctypes stub a maintainer of the reference
 * would add to route that function here.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types.
 *   - unless a function says "host", every data pointer is a DEVICE pointer
 *     owned by the caller; the library allocates nothing persistent.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *     the call returns without synchronising.
 *   - return 0 on success, a negative NH_E_* code on failure (never throws);
 *     nh_last_error() returns a thread-local description of the last failure.
 *   - blocks are "block-major": block b of an (B, N, N) tensor occupies N*N
 *     consecutive elements, rows first.  Pixels / predictions / residuals /
 *     reconstructions are int16, coefficients / levels are int32 (the
 *     reference's dtype contract).
 *   - size is 4, 8, 16 or 32; any other value -> NH_E_SIZE (the reference
 *     raises ValueError("Unsupported transform size"), transform.py:151).
 *   - qp is clamped to [0, 51] exactly like quant.py:35.
 *   - intra modes: 0 = planar, 1 = DC, 2..34 = angular (intra.py:1-17).
 */
#ifndef NH_B200_H
#define NH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NH_OK 0
#define NH_E_SIZE (-1)   /* unsupported block size            */
#define NH_E_ARG (-2)    /* bad argument (null, negative, mode out of range, misaligned) */
#define NH_E_CUDA (-3)   /* CUDA runtime error; see nh_last_error() */
#define NH_E_NOMEM (-4)  /* caller-provided scratch too small  */

#define NH_COST_SAD 0
#define NH_COST_SATD 1

/* ------------------------------------------------------------------ meta */
int nh_version(void);
const char* nh_last_error(void);
/* 1 when a CUDA device of compute capability 10.x is usable by this process. */
int nh_device_ok(void);

/* Tables, computed on the host (no device needed).
 * nano_hevc/transform.py:20-135 (DST4, DCT4/8/16/32) -> out[size*size] */
int nh_get_transform_matrix(int size, int use_dst, int32_t* out_host);
/* nano_hevc/intra.py:24-29 INTRA_PRED_ANGLE[mode-2]; mode 2..34 */
int nh_get_intra_pred_angle(int mode, int* angle_out);
/* nano_hevc/quant.py:25-38 get_qp_params */
int nh_get_qp_params(int qp, int* per_out, int* rem_out);
/* nano_hevc/quant.py:21-22 QUANT_SCALE[rem], DEQUANT_SCALE[rem]; rem 0..5 */
int nh_get_quant_scales(int rem, int* quant_scale_out, int* dequant_scale_out);

/* ------------------------------------------- K4: transforms (batched) */
/* nano_hevc/transform.py:154-196 forward_transform(residual, use_dst).
 * residual: (B,N,N) int16 when residual_is_i32 == 0, int32 otherwise. */
int nh_forward_transform(const void* residual, int residual_is_i32, int32_t* coeff,
                         int64_t n_blocks, int size, int use_dst, void* stream);
/* nano_hevc/transform.py:199-238 inverse_transform(coeff, use_dst). */
int nh_inverse_transform(const int32_t* coeff, int32_t* residual, int64_t n_blocks, int size,
                         int use_dst, void* stream);

/* ---------------------------------------------- K5: quant (batched) */
/* nano_hevc/quant.py:41-79 quantize(coeff, qp, size, is_intra) / :126-137 quantize_block.
 * Like the reference, any size is taken through int(log2(size)); 1..63 is accepted. */
int nh_quantize(const int32_t* coeff, int32_t* level, int64_t n_elems, int qp, int size,
                int is_intra, void* stream);
/* nano_hevc/quant.py:82-123 dequantize(level, qp, size) / :140-150 dequantize_block
 * (`size` is accepted and ignored, as in the reference). */
int nh_dequantize(const int32_t* level, int32_t* coeff, int64_t n_elems, int qp, int size,
                  void* stream);

/* ------------------------------------------ K2: predictors (batched) */
/* nano_hevc/intra.py:46-62 intra_dc_predict(top, left, size); top,left (B,N) */
int nh_intra_dc_predict(const int16_t* top, const int16_t* left, int16_t* pred, int64_t n_blocks,
                        int size, void* stream);
/* The same function for reference arrays of another length than `size`: intra.py:61 adds top.sum() and
 * left.sum() of whatever it is given.  top (B, n_top), left (B, n_left). */
int nh_intra_dc_predict_ragged(const int16_t* top, int n_top, const int16_t* left, int n_left,
                               int16_t* pred, int64_t n_blocks, int size, void* stream);
/* nano_hevc/intra.py:81-113 intra_planar_predict(top, left, top_right, bottom_left, size);
 * top,left (B,N); top_right, bottom_left (B,) */
int nh_intra_planar_predict(const int16_t* top, const int16_t* left, const int16_t* top_right,
                            const int16_t* bottom_left, int16_t* pred, int64_t n_blocks, int size,
                            void* stream);
/* nano_hevc/intra.py:116-207 intra_angular_predict(top, left, top_left, mode, size).
 * top, left: (B, 2N+1) with index 0 = the array's corner slot; top_left (B,).
 * modes: (B,) uint8 per-block modes, or NULL to use `mode` for every block.
 * With allow_dc_planar != 0, mode 1 / 0 select DC / planar computed from
 * top[1..N], left[1..N], TR = top[N+1], BL = left[N+1]. */
int nh_intra_predict_modes(const int16_t* top, const int16_t* left, const int16_t* top_left,
                           const uint8_t* modes, int mode, int allow_dc_planar, int16_t* pred,
                           int64_t n_blocks, int size, void* stream);

/* --------------------------------- K3: residual / reconstruct / clip */
/* nano_hevc/intra.py:65-67 residual_block */
int nh_residual_block(const int16_t* orig, const int16_t* pred, int16_t* residual, int64_t n_elems,
                      void* stream);
/* nano_hevc/intra.py:70-72 reconstruct_block (residual int32 is truncated to int16 first) */
int nh_reconstruct_block(const int16_t* pred, const int32_t* residual, int16_t* out,
                         int64_t n_elems, void* stream);
/* nano_hevc/intra.py:75-78 clip_to_pixel_range */
int nh_clip_to_pixel_range(const int16_t* in, int16_t* out, int64_t n_elems, int bit_depth,
                           void* stream);

/* ------------------------------------------------ K6: fused pipeline */
/* predict -> residual -> forward -> quantize -> dequantize -> inverse ->
 * reconstruct -> clip in one kernel (composition of README.md:55-71 /
 * docs/frames_and_panes.md:319-338).  Any of the four outputs may be NULL.
 *
 * DC / planar from given N-sample references (config 2):
 *   orig (B,N,N); top,left (B,N); top_right,bottom_left (B,);
 *   modes (B,) uint8 with values 0/1, or NULL -> `mode` for all blocks. */
/* Note on streams: the size 4 / 8 kernels hand out their work through a 32-bit counter that belongs
 * to the stream of the call (a slot of a static device array that the last warp of a launch re-arms
 * for the next one).  Calls on one stream, and concurrent calls on different streams, are
 * independent; a slot only moves to another stream after an event recorded behind its last launch
 * has completed (more than 4096 distinct streams per device).  A captured CUDA graph containing such
 * a launch must not be replayed concurrently with itself. */
int nh_fused_pipeline_dcplanar(const int16_t* orig, const int16_t* top, const int16_t* left,
                               const int16_t* top_right, const int16_t* bottom_left,
                               const uint8_t* modes, int mode, int64_t n_blocks, int size, int qp,
                               int is_intra, int use_dst, int bit_depth, int16_t* pred,
                               int32_t* coeff, int32_t* levels, int16_t* recon, void* stream);
/* The three nh_set_*_impl selectors below exist for A/B profiling.  They are PER CALLING THREAD
 * (thread-local): a selection affects the later calls of the thread that made it and nothing else, so
 * the library keeps no mutable state shared between callers.
 * Selects the kernel generation behind nh_fused_pipeline_dcplanar for size 4 / 8:
 * 4 (default: at size 8 the four transform passes run as warp-level f16 tensor-core MMAs, exact
 * for 8-bit samples with an exact fallback otherwise; size 4 uses a rolled one-block-per-lane
 * variant of generation 2), 2 (cp.async
 * prefetch + in-thread butterflies + 32-bit pixel-domain quant with an exact fallback), 3 (2 with
 * TMA tensor-map staging) or 1 (first generation, kept for A/B profiling).  Results are identical. */
int nh_set_fused_impl(int generation);
/* Selects the kernel behind nh_fused_pipeline_dcplanar for size 16 / 32: 2 (default: the four
 * transform passes as warp-level f16 tensor-core MMAs, exact for 8-bit samples, exact CUDA-core
 * fallback otherwise) or 1 (CUDA-core butterflies).  Results are identical. */
int nh_set_rows_impl(int impl);
/* Any of the 35 modes from padded (B, 2N+1) references (K1 convention). */
int nh_fused_pipeline_modes(const int16_t* orig, const int16_t* top, const int16_t* left,
                            const int16_t* top_left, const uint8_t* modes, int mode,
                            int64_t n_blocks, int size, int qp, int is_intra, int use_dst,
                            int bit_depth, int16_t* pred, int32_t* coeff, int32_t* levels,
                            int16_t* recon, void* stream);

/* ---------------------------------------------- K1: reference gather */
/* nano_hevc/block.py:38-55 BlockView.get_top_neighbors / get_left_neighbors /
 * get_top_left_neighbor for every full block of a plane in iterate_blocks
 * order (block.py:68-74), with the 128 substitution at frame edges and
 * replicate-last padding to 2N+1 entries (intra.py:174-178).
 * plane (H, pitch) int16; n_top / n_left = samples requested per side
 * (2N/2N for source neighbours, 2N/N for reconstructed neighbours).
 * top,left (B, 2N+1) [index 0 = corner]; corner (B,). */
int nh_gather_refs(const int16_t* plane, int height, int width, int pitch, int size, int n_top,
                   int n_left, int16_t* top, int16_t* left, int16_t* corner, void* stream);
/* (H,W) plane <-> (B,N,N) block-major, iterate_blocks order, partial blocks skipped. */
int nh_plane_to_blocks(const int16_t* plane, int height, int width, int pitch, int size,
                       int16_t* blocks, void* stream);
int nh_blocks_to_plane(const int16_t* blocks, int height, int width, int pitch, int size,
                       int16_t* plane, void* stream);

/* ---------------------------------- K7 / K8: frame coders with search */
/* Exhaustive 35-mode search (candidate order 1,0,2..34; first strict minimum,
 * i.e. DC beats planar on ties -- __main__.py:96 -- then the lowest angular
 * mode) with SAD (metrics.py:24-26) or SATD (sum of satd_4x4, metrics.py:29-43)
 * cost, then the winner through the K6 chain with use_dst = (size == 4).
 *
 * recon_neighbours == 0 (config 3): references come from the SOURCE plane,
 *   n_top = n_left = 2N; every block independent.
 * recon_neighbours == 1 (config 5): references come from the RECON plane
 *   (zero-initialised by this call), n_top = 2N, n_left = N; blocks are coded
 *   in an anti-diagonal wavefront that honours the raster-order dependencies.
 *
 * src (H, pitch) int16.  Outputs, block-major in raster block order (any may be
 * NULL except recon_plane when recon_neighbours == 1):
 *   modes (B,) u8; costs (B,) i32; pred (B,N,N) i16; coeff, levels (B,N,N) i32;
 *   recon_plane (H, pitch) i16 -- uncovered rows/columns are set to 0.
 * scratch: device scratch of nh_encode_frame_scratch_bytes(height, width, size) bytes
 *   (only used when recon_neighbours == 1; may be NULL otherwise). */
int64_t nh_encode_frame_scratch_bytes(int height, int width, int size);
/* Selects how recon_neighbours == 0 runs on 8-bit content: 2 (default: a search kernel decides the
 * modes -- every lane of a warp evaluates the same candidate mode on its own strip of pixels --
 * and a second kernel codes the winners; needs the modes tensor, pitch % 4 == 0 and an 8-byte
 * aligned plane, otherwise 1 is used) or 1 (search and winner pipeline in one kernel).  At N >= 8
 * there are two search kernels (strips of 4 scan lines per lane; all lanes on the same scan line,
 * which needs pitch % 8 == 0 and a 16-byte aligned plane): 2 picks per call by size and cost kind,
 * 3 / 4 force the second / the first, 5 the fraction-major kernel (N = 16 / 32), 6 the SATD kernel
 * with the 4x4 Hadamard transforms on the tensor cores (SATD; N >= 8 needs pitch % 8 == 0 and a 16-byte
 * aligned plane; what 2 picks for SATD) -- profiling, tests.  Results are identical.  The setting is per
 * calling thread. */
int nh_set_search_impl(int impl);
/* Selects how recon_neighbours == 1 (the anti-diagonal wavefront over block rows) is laid out, for the calling
 * thread.  warps: 0 (default) = pick per call from the block rows in flight (rows of a frame x frames of the call),
 * 1 / 2 / 4 / 8 / 12 = warps per block row at N = 16 / 32 (more warps: shorter dependent block time, fewer rows
 * resident; 12 exists at N = 32 only and means 8 at N = 16),
 * 1 / 4 = the one-warp / four-warp kernel at N = 4; ignored at N = 8 (always four warps per row on 8-bit planes).
 * build: 0 (default) = pick per call, 1 = latency build (one or two frames), 2 = throughput build (more CTAs per SM),
 * 3 = throughput build at the highest occupancy (N = 8; elsewhere as 2); N = 4 / 8 only.  Results are identical. */
int nh_set_wave_impl(int warps, int build);
int nh_encode_frame(const int16_t* src, int height, int width, int pitch, int size, int cost_kind,
                    int qp, int recon_neighbours, int bit_depth, uint8_t* modes, int32_t* costs,
                    int16_t* pred, int32_t* coeff, int32_t* levels, int16_t* recon_plane,
                    void* scratch, int64_t scratch_bytes, void* stream);

/* The same coders over a BATCH of frames in one call (BASELINE configs 3 and 5: "32 x 4K", "8-frame 4K
 * batch"): frame f is the (height, pitch) plane at src + f * frame_stride (frame_stride >= height * pitch,
 * in samples; recon_planes uses the same strides) and every block-major output holds frame f at block
 * offset f * (height / size) * (width / size).  Frames are independent -- each is coded exactly as
 * nh_encode_frame would code it -- but one launch covers all of them: with recon_neighbours != 0 the block
 * rows of all frames share the wavefront scheduler (tickets interleave the frames), so F frames fill the
 * GPU where one frame occupies a few hundred warps.
 * stats: optional (F, 4) int64 device tensor, written by this call:
 *   { sum (src - recon)^2 over the whole plane (numerator of metrics.py:7-21; uncovered rows count as in
 *     __main__.py:135-137), height * width, sum of the winners' costs, number of non-zero levels };
 *   needs recon_planes, and costs / levels for the last two entries (0 otherwise).
 * scratch: nh_encode_frames_scratch_bytes(n_frames, height, width, size) bytes (recon_neighbours != 0 only). */
int64_t nh_encode_frames_scratch_bytes(int n_frames, int height, int width, int size);
int nh_encode_frames(const int16_t* src, int n_frames, int64_t frame_stride, int height, int width,
                     int pitch, int size, int cost_kind, int qp, int recon_neighbours, int bit_depth,
                     uint8_t* modes, int32_t* costs, int16_t* pred, int32_t* coeff, int32_t* levels,
                     int16_t* recon_planes, int64_t* stats, void* scratch, int64_t scratch_bytes,
                     void* stream);

/* -------------------------------------------------- K9: reductions */
/* Integer numerators of nano_hevc/metrics.py: out[0] = sum (a-b)^2  (mse/psnr, :7-21),
 * out[1] = sum |a-b| (sad, :24-26).  a, b int16, n_elems elements; out: 2 x int64 on the
 * device, zeroed by this call.  PSNR is finished on the host in float64. */
int nh_reduce_sse_sad(const int16_t* a, const int16_t* b, int64_t n_elems, int64_t* out,
                      void* stream);
/* Same over a (height, width) window of two pitched planes. */
int nh_reduce_sse_sad_2d(const int16_t* a, int pitch_a, const int16_t* b, int pitch_b, int height,
                         int width, int64_t* out, void* stream);
/* Wide-input reductions behind the per-block metric wrappers: the reference widens before it reduces
 * (metrics.py:9 float64, :26 / :33 int32, :48 int64), so uint16 samples and the int32 output of
 * inverse_transform must not be narrowed to int16.  a, b int32 (b may be NULL = zeros), n elements.
 *   out[0] = sum (a-b)^2, difference in int64, accumulated modulo 2^64   (residual_energy, metrics.py:46-48)
 *   out[1] = sum |a-b| with the int32 wrap-around of metrics.py:26, summed in int64   (sad)
 *   *fsum  = sum of float64 (a-b)^2                                       (mse for sums beyond 2^53)
 * out: 2 x int64, fsum: 1 x double, device, zeroed by this call. */
int nh_reduce_metrics_i32(const int32_t* a, const int32_t* b, int64_t n_elems, int64_t* out, double* fsum,
                          void* stream);
/* metrics.py:7-10 mse for float64 inputs: *fsum = sum (a-b)^2 in float64 (summation order differs from
 * numpy's pairwise sum: equal to ~1e-15 relative). */
int nh_reduce_sse_f64(const double* a, const double* b, int64_t n_elems, double* fsum, void* stream);
/* metrics.py:29-43 satd_4x4 on int32 inputs with the reference's int32 arithmetic: a, b (B,4,4) int32,
 * out (B,) int64. */
int nh_satd_4x4_i32(const int32_t* a, const int32_t* b, int64_t n_blocks, int64_t* out, void* stream);
/* Per-block costs of (B,N,N) pairs: sad (B,) i32 and satd (B,) i32 (sum of satd_4x4 over the
 * 4x4 sub-blocks, metrics.py:29-43), energy (B,) i64 = sum (a-b)^2 (residual_energy, :46-48).
 * Any output may be NULL. */
int nh_block_costs(const int16_t* a, const int16_t* b, int64_t n_blocks, int size, int32_t* sad,
                   int32_t* satd, int64_t* energy, void* stream);
/* Level statistics: out[0] = number of non-zero levels (quant.py:171-173 count_nonzero). */
int nh_count_nonzero(const int32_t* levels, int64_t n_elems, int64_t* out, void* stream);

/* nano_hevc/quant.py:153-168 estimate_bits: *nnz_out = number of non-zero levels, *sum_log2_out =
 * sum over the levels of log2(|level| + 1) in float64 (device scalars, zeroed by this call).  The
 * estimate is int(sum_log2 + 2 * nnz), finished on the host. */
int nh_level_stats(const int32_t* levels, int64_t n_elems, int64_t* nnz_out, double* sum_log2_out,
                   void* stream);

/* ------------------------------------- host-buffer entry point (e2e) */
/* Same computation as nh_fused_pipeline_dcplanar but every pointer is a HOST
 * pointer (pinned memory recommended).  The library copies the inputs to the
 * current device in chunks, runs the kernel and copies the requested outputs
 * back, overlapping H2D, compute and D2H on internal streams, and returns
 * after the last byte has landed.  device_scratch is a caller-provided device
 * buffer of at least nh_host_pipeline_scratch_bytes(size, chunk_blocks) bytes. */
int64_t nh_host_pipeline_scratch_bytes(int size, int64_t chunk_blocks);
/* Bytes the most recent nh_host_pipeline_dcplanar call moved over PCIe in each direction (the
 * output wire format is compact and data dependent: int8 coefficients with int16 exception
 * segments, all-zero level segments elided; see csrc/nh_host.cu).  Either pointer may be NULL. */
int nh_host_pipeline_last_transfer(int64_t* h2d_bytes, int64_t* d2h_bytes);
int nh_host_pipeline_dcplanar(const int16_t* orig, const int16_t* top, const int16_t* left,
                              const int16_t* top_right, const int16_t* bottom_left,
                              const uint8_t* modes, int mode, int64_t n_blocks, int size, int qp,
                              int is_intra, int use_dst, int bit_depth, int16_t* pred,
                              int32_t* coeff, int32_t* levels, int16_t* recon,
                              void* device_scratch, int64_t scratch_bytes, int64_t chunk_blocks);

/* nh_host_pipeline_dcplanar with coefficients and levels delivered as int16 (an option next to the reference's
 * int32 dtypes; same arguments otherwise, same device scratch): results go by DMA straight into the caller's
 * arrays, no host thread touches them.  Valid while every block stays in the pixel domain (samples in [0, 4095]:
 * |coeff| <= 32394, |level| <= 13600); otherwise the call fails with NH_E_ARG after its transfers have drained. */
int nh_host_pipeline_dcplanar_i16(const int16_t* orig, const int16_t* top, const int16_t* left,
                                  const int16_t* top_right, const int16_t* bottom_left,
                                  const uint8_t* modes, int mode, int64_t n_blocks, int size, int qp,
                                  int is_intra, int use_dst, int bit_depth, int16_t* pred,
                                  int16_t* coeff16, int16_t* levels16, int16_t* recon,
                                  void* device_scratch, int64_t scratch_bytes, int64_t chunk_blocks);

/* ------------------------------------- host-buffer entry of the frame coders (configs 3 / 5 end to end) */
/* nh_encode_frames for frames and results in HOST memory (docs/frames_and_panes.md:319-344 with numpy arrays
 * on both sides): src is n_frames contiguous (height, width) int16 planes; every output is optional (NULL = not
 * delivered) and has the layout of nh_encode_frames with pitch = width; stats is a HOST (n_frames, 4) int64
 * array.  The batch is cut into chunks of frames_per_chunk frames that rotate over three internal streams
 * (upload, kernels and download of neighbouring chunks overlap); the call returns when every result is in
 * place.  Page-locked caller buffers give asynchronous copies at link speed; pageable ones work.
 * device_scratch: nh_host_encode_frames_scratch_bytes(frames_per_chunk, height, width, size, recon_neighbours)
 *   bytes of device memory, 256-byte aligned, idle for the duration of the call.
 * nh_host_encode_frames_last_transfer: bytes the last call moved over PCIe in each direction. */
int64_t nh_host_encode_frames_scratch_bytes(int frames_per_chunk, int height, int width, int size,
                                            int recon_neighbours);
int nh_host_encode_frames_last_transfer(int64_t* h2d_bytes, int64_t* d2h_bytes);
int nh_host_encode_frames(const int16_t* src, int n_frames, int height, int width, int size, int cost_kind,
                          int qp, int recon_neighbours, int bit_depth, uint8_t* modes, int32_t* costs,
                          int16_t* pred, int32_t* coeff, int32_t* levels, int16_t* recon_planes,
                          int64_t* stats, int frames_per_chunk, void* device_scratch, int64_t scratch_bytes);

/* ------------------------------------- device-side frame containers */
/* Sample conversion for planes held on the device (nano_hevc/frame.py): uint8 -> int16 zero-extends
 * (frame.py:45-51, Plane.from_buffer followed by the coder's astype(int16)); int16 -> uint8 keeps the
 * low 8 bits, which is what numpy's astype(np.uint8) in Frame.to_yuv420p / PackedFrame.to_yuv420p
 * does (frame.py:107-111, :172-178).  n samples, device pointers. */
int nh_convert_u8_to_i16(const uint8_t* src, int16_t* dst, int64_t n, void* stream);
int nh_convert_i16_to_u8(const int16_t* src, uint8_t* dst, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NH_B200_H */
