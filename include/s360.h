/* s360.h — C ABI of the B200-native Surround360 stereo-panorama render path.
 *
 * Drop-in boundary for surround360_render's per-frame renderer
 * (surround360_render/source/test/TestRenderStereoPanorama.cpp). The reference has no
 * FFI layer; these entry points are what a C++ shim placed under the reference's
 * in-process operator API would bind (see INTEGRATION.md for the ~30-line adapters):
 *
 *   s360_flow_*            <- OpticalFlowInterface::computeOpticalFlow
 *                             (source/optical_flow/OpticalFlowInterface.h:34-41) as created by
 *                             makeOpticalFlowByName (source/optical_flow/OpticalFlowFactory.h:23-64)
 *   s360_novel_view_chunk  <- NovelViewGenerator::prepare + combineLazyNovelViews
 *                             (source/optical_flow/NovelView.h:79-138, NovelView.cpp:226-299)
 *   s360_project_spherical <- bicubicRemapToSpherical (source/render/ImageWarper.h:59-66) with
 *                             projectSideToSpherical's alpha ramp (TRSP:99-135)
 *   s360_flatten_deghost   <- flattenLayersDeghostPreferBase (source/util/CvUtil.h:72-74)
 *   s360_feather_alpha     <- featherAlphaChannel (source/util/CvUtil.cpp:140-157)
 *   s360_sharpen           <- iirLowPass + sharpenWithIirLowPass (source/util/Filter.h:39-137)
 *   s360_cubemap           <- convertSphericalToCubemapBicubicRemap + stackOutputCubemapFaces
 *                             (source/render/ImageWarper.h:53-57, source/util/CvUtil.cpp:117-138)
 *   s360_resize_cubic      <- cv::resize(INTER_CUBIC) at TRSP:943-956
 *   s360_render_frame      <- renderStereoPanorama() (TRSP:716-972) minus file I/O
 *   s360_hyper_preview     <- PreviewRenderer::render (source/test/TestHyperPreview.cpp:77-185)
 *   s360_isp_*             <- CameraIsp (source/camera_isp/CameraIsp.h:44-1301) as driven by Raw2Rgb's CPU branch
 *                             (source/camera_isp/Raw2Rgb.cpp:445-456)
 *   s360_rig_*             <- Camera::loadRig / RigDescription (source/render/Camera.cpp:243-254,
 *                             source/render/RigDescription.cpp:18-78)
 *
 * Conventions: plain pointers and sizes; images are tightly packed row-major unless a
 * stride is given (the *_strided variants); BGRA8 / BGR8 / float2 flow exactly as the reference's cv::Mat types
 * (CV_8UC4 / CV_8UC3 / CV_32FC2). Every function returns 0 on success, non-zero on error
 * (no exceptions cross the ABI); s360_last_error() returns the message for the calling
 * thread. A context is bound to one CUDA device and must not be shared between host
 * threads concurrently. Buffers are owned by the caller. `_dev` variants take device
 * pointers (same layout) and enqueue on the context's stream without synchronising.
 * There is no CPU fallback: if the CUDA device or kernels are unavailable the call fails.
 */
#ifndef S360_H
#define S360_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct s360_ctx s360_ctx;

/* OpticalFlowInterface::DirectionHint (OpticalFlowInterface.h:24) — same numeric order */
typedef enum { S360_HINT_UNKNOWN = 0, S360_HINT_RIGHT = 1, S360_HINT_DOWN = 2, S360_HINT_LEFT = 3, S360_HINT_UP = 4 } s360_hint;

/* Pixel-affecting gflags of TRSP:44-70 (same names, same defaults via s360_params_default) */
typedef struct {
  double interpupilary_dist;      /* 6.4 */
  double sharpening;              /* 0.0 */
  double zero_parallax_dist;      /* 10000.0 */
  int side_alpha_feather_size;    /* 100 */
  int std_alpha_feather_size;     /* 31 */
  int enable_top;                 /* 0 */
  int enable_bottom;              /* 0 */
  int eqr_width;                  /* 256 */
  int eqr_height;                 /* 128 */
  int final_eqr_width;            /* 3480 */
  int final_eqr_height;           /* 960 */
  int cubemap_width;              /* 1536 */
  int cubemap_height;             /* 1536 */
  int want_cubemap;               /* output_cubemap_path non-empty */
  int cubemap_photo;              /* cubemap_format == "photo" (default "video") */
  char side_flow_alg[32];         /* "pixflow_low" */
  char polar_flow_alg[32];        /* "pixflow_low" */
  int enable_pole_removal;        /* 0; needs s360_set_pole_masks */
  int reserved0;
  char poleremoval_flow_alg[32];  /* "pixflow_low" */
} s360_params;

const char* s360_last_error(void);
const char* s360_version(void);
void s360_params_default(s360_params* p);

/* ---- context ---- */
int s360_ctx_create(int device, s360_ctx** out);
void s360_ctx_destroy(s360_ctx* ctx);
int s360_ctx_synchronize(s360_ctx* ctx);
/* cudaStream_t of the context, as void* (for callers that enqueue their own work) */
void* s360_ctx_stream(s360_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
unsigned long long s360_ctx_launch_count(s360_ctx* ctx);
/* keep copies of the pre-sharpen eye images for s360_frame_intermediate("eyeL"/"eyeR") */
int s360_ctx_set_keep_intermediates(s360_ctx* ctx, int on);

/* Profiling of the dominant kernel (the PixFlow wavefront sweep): when enabled, CUDA events
 * are recorded on the context stream around every sweep launch of s360_render_frame*.
 * s360_profile_sweep synchronises the stream and returns the accumulated device time (ms),
 * number of launches, level-pixels processed (flows x width x height per launch) and
 * wavefront steps since s360_profile_enable(ctx, 1). on = 2 additionally brackets EVERY kernel launch of the frame with events
 * (s360_profile_kernels); that serialises the frame's launches much more than level 1 does. */
int s360_profile_enable(s360_ctx* ctx, int on);
/* Whole-frame CUDA-graph replay (default on). Profiling, keep_intermediates and the stage timers
 * need the plain launch path and switch it off implicitly for the frames they cover. */
int s360_ctx_set_graphs(s360_ctx* ctx, int on);
/* Latency mode (default off): the optical-flow sweeps of this context run the kernel variant that shortens ONE frame (four lanes per
 * pixel, inputs and the own-flow footprint of each pixel staged through shared memory ahead of the dependent chain: 219 instead of 266 ms
 * per 8K frame rendered alone on a B200) at the price of issue slots and shared memory that lower the frame rate when many contexts
 * render at once. Use it for previews / single-frame renders; leave it off for batch throughput. Results are bit-identical in both modes. */
int s360_ctx_set_latency_mode(s360_ctx* ctx, int on);
/* Video mode (--prev_frame_data_dir != NONE, scripts/batch_process_video.py:137,157-158): when on, every frame
 * rendered on this context regularises its flows against the previous frame rendered on it (PixFlow.h:101-118,
 * 147-153, 165-167); the previous flows / overlap images / extended fisheye images stay on the device instead of
 * going through .bin and PNG files (TRSP:200-255, 413-452). The first frame after enabling or after
 * s360_ctx_reset_temporal renders like NONE. */
int s360_ctx_set_temporal(s360_ctx* ctx, int on);
int s360_ctx_reset_temporal(s360_ctx* ctx);
/* Video mode through files, as the reference hands a frame's state to the next process (TRSP:200-255, 413-452):
 * with temporal mode on, s360_frame_prepare builds the frame plan for (params, image size) without rendering, so that
 * s360_temporal_sizes can report the sizes of the per-pair overlap images / side flows (overlap_w x overlap_h) and of the
 * per-pole-slot extended images / pole flows (ext_w x ext_h). s360_temporal_import_side loads what the reference reads
 * from flow/<prev>/flowLtoR_<pair>.bin, flowRtoL_<pair>.bin and debug/<prev>/flow_images/overlap_<pair>_{L,R}.png
 * (BGRA8); s360_temporal_import_pole what it reads from flow/<prev>/flow_<slot>.bin and
 * extendedFisheyeSpherical_<slot>.png (slot 0..3 = top_left, top_right, bottom_left, bottom_right; the extended side
 * image of the previous frame is read by the reference but not used by PixFlow, PixFlow.h:101-118).
 * s360_temporal_commit marks the imported state as "the previous frame" of the next s360_render_frame.
 * The matching exports are the s360_frame_intermediate names "flowLtoR", "flowRtoL", "overlapL", "overlapR" (idx = pair)
 * and "poleFlow", "extSide", "extFish" (idx = slot). */
int s360_frame_prepare(s360_ctx* ctx, const s360_params* p, int img_w, int img_h);
int s360_temporal_sizes(s360_ctx* ctx, int* overlap_w, int* overlap_h, int* ext_w, int* ext_h);
int s360_temporal_import_side(s360_ctx* ctx, int pair, const float* flow_l_to_r, const float* flow_r_to_l, const uint8_t* overlap_l_bgra,
                              const uint8_t* overlap_r_bgra);
int s360_temporal_import_pole(s360_ctx* ctx, int slot, const float* flow, const uint8_t* ext_fisheye_bgra);
int s360_temporal_commit(s360_ctx* ctx);
int s360_profile_sweep(s360_ctx* ctx, double* ms, unsigned long long* launches, double* level_pixels, double* steps);
/* sweep launches (since s360_profile_enable) that the generic kernel had to repeat because an operand left the range of the
 * fast kernel's written-out div/sqrt sequences; valid after s360_profile_sweep */
int s360_profile_sweep_fallbacks(s360_ctx* ctx, unsigned long long* fallbacks);
/* Per-kernel table of everything launched by s360_render_frame[_dev] since s360_profile_enable(ctx, 2): a JSON array of
 * {"kernel", "ms" (CUDA events on the launching stream, summed), "algorithmic_bytes" (compulsory reads + writes of those launches),
 * "launches"}. Synchronises the stream. *needed receives the size of the text incl. the terminator; json may be NULL to query it. */
int s360_profile_kernels(s360_ctx* ctx, char* json, size_t cap, size_t* needed);

/* ---- rig (RIG_JSON.md format) ---- */
int s360_rig_load_file(s360_ctx* ctx, const char* json_path);
int s360_rig_load_json(s360_ctx* ctx, const char* json_text);
int s360_rig_num_cameras(s360_ctx* ctx);
int s360_rig_num_side(s360_ctx* ctx);
/* which: 0 = top, 1 = bottom, 2 = secondary bottom; returns index into the rig, <0 on error */
int s360_rig_special_camera(s360_ctx* ctx, int which);
/* index into the rig of side camera `side_idx` */
int s360_rig_side_index(s360_ctx* ctx, int side_idx);
const char* s360_rig_camera_id(s360_ctx* ctx, int cam);
/* Camera::pixel (Camera.h:133-140), host evaluation in double */
int s360_camera_pixel(s360_ctx* ctx, int cam, const double xyz[3], double out_xy[2]);
/* Compact binary form of the parsed rig + projection tables, for the one-off NCCL broadcast:
 * rank 0 serialises, every rank deserialises. */
int s360_rig_serialize(s360_ctx* ctx, void* buf, size_t cap, size_t* used);
int s360_rig_deserialize(s360_ctx* ctx, const void* buf, size_t len);

/* Pole masks of --bottom_pole_masks_dir (<dir>/<bottomCamId>.png and <dir>/<bottomCam2Id>.png loaded as
 * BGR8, TRSP:569-597, PoleRemoval.cpp:58-67): pixels equal to pure red (0,0,255) are cut out of the alpha
 * channel. Host pointers, both w x h (the camera image size); copied to the device. */
int s360_set_pole_masks(s360_ctx* ctx, const uint8_t* mask_bottom_bgr, const uint8_t* mask_bottom2_bgr, int w, int h);

/* ---- geometry derived from rig + params (TRSP:155-162,309-316,340-348) ---- */
typedef struct {
  int cam_w, cam_h, overlap_w, num_novel;
  float h_radians, v_radians, fov_h_radians, slab_disp, shift_px;
  int pole_h;   /* rows of the top/bottom spherical images (TRSP:607-610) */
} s360_geometry;
int s360_geometry_query(s360_ctx* ctx, const s360_params* p, s360_geometry* out);

/* ---- operators (host-pointer variants; synchronous) ---- */

/* bicubicRemapToSpherical (+ optional side alpha ramp). src: BGR8 (src_channels 3) or BGRA8
 * (4). dst: BGRA8 dst_w x dst_h. feather > 0 applies projectSideToSpherical's ramp to the
 * source rows (TRSP:116-125). */
int s360_project_spherical(s360_ctx* ctx, int cam, const uint8_t* src, int src_w, int src_h, int src_channels,
                           float left_angle, float right_angle, float top_angle, float bottom_angle, int feather,
                           uint8_t* dst_bgra, int dst_w, int dst_h);

/* computeOpticalFlow(I0BGRA, I1BGRA, prevFlow, prevI0BGRA, prevI1BGRA, flow, hint).
 * prev_* may all be NULL (no temporal regularisation). flow: w*h float2. */
int s360_flow_compute(s360_ctx* ctx, const char* alg_name, const uint8_t* i0_bgra, const uint8_t* i1_bgra, int w, int h,
                      const float* prev_flow, const uint8_t* prev_i0_bgra, const uint8_t* prev_i1_bgra, float* flow_out,
                      s360_hint hint);
/* Batched variant: n independent pairs of identical size, solved concurrently. */
int s360_flow_compute_batch(s360_ctx* ctx, const char* alg_name, int n, const uint8_t* const* i0_bgra,
                            const uint8_t* const* i1_bgra, int w, int h, float* const* flow_out, const s360_hint* hints);

/* Stride-aware variant for non-contiguous cv::Mat views (the reference's overlap images are ROI crops, TRSP:196-198): every
 * *_stride is the distance between rows in BYTES (0 = tightly packed). prev_i0 is not needed (the reference resizes it and never
 * reads it, PixFlow.h:106). */
int s360_flow_compute_strided(s360_ctx* ctx, const char* alg_name, const uint8_t* i0_bgra, size_t i0_stride, const uint8_t* i1_bgra,
                              size_t i1_stride, int w, int h, const float* prev_flow, size_t prev_flow_stride, const uint8_t* prev_i1_bgra,
                              size_t prev_i1_stride, float* flow_out, size_t flow_out_stride, s360_hint hint);
/* Device-pointer variant: BGRA8 images resident in HBM (row strides in bytes, multiples of 4; 0 = packed), flow written densely
 * (w*h float2) to flow_out_dev. Enqueued on the context stream without synchronising; the solver for (algorithm, size) is
 * created on first use and kept by the context. */
int s360_flow_compute_dev(s360_ctx* ctx, const char* alg_name, const uint8_t* i0_bgra_dev, size_t i0_stride, const uint8_t* i1_bgra_dev,
                          size_t i1_stride, int w, int h, float* flow_out_dev, s360_hint hint);

/* NovelViewGeneratorLazyFlow::combineLazyNovelViews for one camera pair with the warp buffer
 * of renderStereoPanoramaChunksThread (TRSP:259-292). Outputs two BGRA8 chunks buf_w x h. */
int s360_novel_view_chunk(s360_ctx* ctx, const uint8_t* overlap_l, const uint8_t* overlap_r, const float* flow_l_to_r,
                          const float* flow_r_to_l, int w, int h, int buf_w, int cam_w, int num_novel, float slab_disp,
                          uint8_t* chunk_l, uint8_t* chunk_r);

/* Device-pointer variant (overlap images share one row stride in bytes, 0 = packed; flows and chunks dense); no synchronisation. */
int s360_novel_view_chunk_dev(s360_ctx* ctx, const uint8_t* overlap_l_dev, const uint8_t* overlap_r_dev, size_t overlap_stride,
                              const float* flow_l_to_r_dev, const float* flow_r_to_l_dev, int w, int h, int buf_w, int cam_w, int num_novel,
                              float slab_disp, uint8_t* chunk_l_dev, uint8_t* chunk_r_dev);
/* s360_project_spherical with row strides in bytes for source and destination (0 = packed) */
int s360_project_spherical_strided(s360_ctx* ctx, int cam, const uint8_t* src, size_t src_stride, int src_w, int src_h, int src_channels,
                                   float left_angle, float right_angle, float top_angle, float bottom_angle, int feather,
                                   uint8_t* dst_bgra, size_t dst_stride, int dst_w, int dst_h);

int s360_flatten_deghost(s360_ctx* ctx, const uint8_t* base_bgra, const uint8_t* top_bgra, int w, int h, uint8_t* out_bgra);
int s360_feather_alpha(s360_ctx* ctx, const uint8_t* src_bgra, int w, int h, int erode_size, uint8_t* out_bgra);
/* offsetHorizontalWrap (CvUtil.cpp:93-115), channels 3 or 4 */
int s360_offset_wrap(s360_ctx* ctx, const uint8_t* src, int w, int h, int channels, float offset, uint8_t* out);
/* sharpenThread (TRSP:688-696): in-place on BGR8 */
int s360_sharpen(s360_ctx* ctx, uint8_t* bgr, int w, int h, double sharpening);
/* both stages of TRSP:917-935 for one eye; out = stacked faces (video: 3*fw x 2*fh, photo: fw x 6*fh) */
int s360_cubemap(s360_ctx* ctx, const uint8_t* eye_bgr, int w, int h, int face_w, int face_h, int photo, uint8_t* out_bgr);
int s360_resize_cubic(s360_ctx* ctx, const uint8_t* src, int w, int h, int channels, int dst_w, int dst_h, uint8_t* dst);
/* poleToSideFlowThread (TRSP:388-561): side BGRA8 sw x sh (padded eye panorama), fisheye
 * BGRA8 fw x fh; warped_out BGRA8 sw x sh; flow_out (optional) int(fw*1.2) x fh float2 */
int s360_pole_to_side(s360_ctx* ctx, const s360_params* p, const uint8_t* side_bgra, int sw, int sh,
                      const uint8_t* fisheye_bgra, int fw, int fh, uint8_t* warped_out, float* flow_out);

/* TestHyperPreview (source/test/TestHyperPreview.cpp:77-185): the fast mono preview straight from the raw Bayer frames (uint16,
 * img_w x img_h each) of the top, bottom and secondary-bottom cameras of the loaded rig: 2x2 demosaic, alpha fades, cubic
 * projection through the precomputed equirect->camera warp (built on the first call for a size, then resident), softmax flatten,
 * x 255/65535. out_bgr: eqr_h x eqr_w x 3 float32, the image the reference hands to imwrite. Bit-identical to the reference build
 * for gamma == 1 (its default; the blend's expf is an exact port of the host's); gamma != 1 goes through the device's powf (2e-4). */
int s360_hyper_preview(s360_ctx* ctx, const uint16_t* raw_top, const uint16_t* raw_bottom, const uint16_t* raw_bottom2, int img_w, int img_h,
                       int eqr_w, int eqr_h, double gamma, double softmax_coef, float* out_bgr);

/* ---- soft ISP: raw Bayer -> RGB, the step in front of the render path ----
 * s360_isp mirrors the reference's CameraIsp object (source/camera_isp/CameraIsp.h): constructed from the ISP configuration JSON
 * (res/config/isp/<name>.json) and the output depth (8 or 16), configured through the same setters, then fed frames. The pipeline is
 * CameraIsp::executePipeline (:1262-1272): black level, anti-vignetting, white balance, clamp and stretch, demosaic, colour matrix +
 * tone curve, IIR unsharp mask; bit-identical to the reference build. demosaic_filter: 0 bilinear, 2 edge aware (the reference's
 * default); 1 (frequency domain, cv::dct) is rejected. A configuration with stuckPixelRadius > 0 and stuckPixelThreshold <= 1 is
 * rejected (sequential scan; with the shipped thresholds the reference's stuck-pixel loop never runs). */
typedef struct s360_isp s360_isp;
int s360_isp_create(const char* isp_json, int output_bpp, s360_isp** out);      /* CameraIsp(jsonInput, outputBpp), :428 */
void s360_isp_destroy(s360_isp* isp);
int s360_isp_set_bits_per_pixel(s360_isp* isp, int bits);                        /* setBitsPerPixel, :860 */
int s360_isp_set_demosaic_filter(s360_isp* isp, int filter);                     /* setDemosaicFilter, :968 */
int s360_isp_set_resize(s360_isp* isp, int resize);                              /* setResize, :978 (1, 2, 4, 8) */
int s360_isp_set_tone_map(s360_isp* isp, int enable);                            /* enableToneMap / disableToneMap, :958-966 */
int s360_isp_add_black_level_offset(s360_isp* isp, int offset);                  /* addBlackLevelOffset, :884 */
/* The tables setup() derives, for an image of w x h after the resize (host only, no device needed): compositeCCM (9 floats, row
 * major), toneCurveLut (4096 x 3), vignette gain of every column (w x 3) and row (h x 3). Any output may be NULL. */
int s360_isp_tables(s360_isp* isp, int w, int h, float* ccm9, float* lut, float* vig_h, float* vig_v);
/* loadImage + getImage (:831-857, :1275-1299). raw: in_w x in_h samples of depth_bits (8 or 16), stride_bytes between rows (0 =
 * tightly packed); out: (in_h / resize) x (in_w / resize) x 3 of output_bpp bits, tightly packed, BGR when swizzle != 0 (the
 * reference's default) else RGB. Host buffers; the copies and kernels run on the context's stream. */
int s360_isp_process(s360_ctx* ctx, s360_isp* isp, const void* raw, int in_w, int in_h, int depth_bits, size_t stride_bytes, void* out, int swizzle);
/* Same on DEVICE buffers, asynchronous on the context's stream: the 8-bit BGR output can be handed to s360_render_frame_dev without
 * the 16-bit PNG round trip of the reference (Unpacker.cpp:187-190 -> RigDescription.cpp:97-101). */
int s360_isp_process_dev(s360_ctx* ctx, s360_isp* isp, const void* raw_dev, int in_w, int in_h, int depth_bits, size_t stride_bytes, void* out_dev,
                         int swizzle);

/* The renderer's input straight from a raw frame: an ISP created with 16 output bits (as Unpacker.cpp:174-176 does), its BGR result
 * narrowed to 8 bits the way the renderer's cv::imread narrows Unpacker's 16-bit PNG (high byte). out_bgr8_dev: (in_h / resize) x
 * (in_w / resize) x 3 uint8 on the device, ready for s360_render_frame_dev. Asynchronous on the context's stream. */
int s360_isp_process_frame_dev(s360_ctx* ctx, s360_isp* isp, const void* raw_dev, int in_w, int in_h, int depth_bits, size_t stride_bytes,
                               uint8_t* out_bgr8_dev);

/* The same for the n cameras of a frame (n <= 32, one size) through shared launches: pointer arrays on the HOST, images on the device. */
int s360_isp_process_frames_dev(s360_ctx* ctx, s360_isp* isp, int n, const void* const* raw_dev, int in_w, int in_h, int depth_bits,
                                size_t stride_bytes, uint8_t* const* out_bgr8_dev);

/* ---- footage container and raw-sample conversion: the Unpacker's front end ----
 * s360_footage mirrors BinaryFootageFile (source/camera_isp/BinaryFootageFile.h:16-62): the capture software's ".bin" file, a 32-byte
 * header and, from offset 4096, frame f of camera c at (number_of_cameras * f + c) * frame_bytes. Host I/O (read-only mapping). */
typedef struct s360_footage s360_footage;
typedef struct {   /* BinaryFootageFile::MetadataHeader, BinaryFootageFile.h:18-27 */
  uint32_t magic, timestamp, file_index, file_count, width, height, bits_per_pixel, number_of_cameras;
} s360_footage_header;
int s360_footage_open(const char* path, s360_footage** out);                        /* BinaryFootageFile(path) + open(), .cpp:30-101 */
void s360_footage_close(s360_footage* f);
/* getMetadata / getNumberOfFrames (.cpp:160-168) / getFrameSize (.cpp:170-172); any output may be NULL */
int s360_footage_info(s360_footage* f, s360_footage_header* header, uint64_t* frames, uint64_t* frame_bytes);
/* getFrame (.cpp:143-154): pointer into the mapping, valid until s360_footage_close; error when the camera or frame is outside the file */
int s360_footage_frame(s360_footage* f, uint32_t frame, uint32_t camera, const uint8_t** data);

/* RawConverter::convert8Frame / convert12Frame (source/camera_isp/RawConverter.cpp:14-59): packed footage samples (bits 8 or 12, dense:
 * w * h * bits / 8 bytes) -> w x h uint16 (8-bit: v * 0x101; 12-bit: unswizzled and replicated to 16 bits). Host buffers. */
int s360_raw_convert(s360_ctx* ctx, const void* packed, int bits, int w, int h, uint16_t* out);
/* n frames of one size, device buffers (pointer arrays on the HOST), asynchronous on the context's stream */
int s360_raw_convert_dev(s360_ctx* ctx, int n, const void* const* packed_dev, int bits, int w, int h, uint16_t* const* out_dev);
/* Unpacker.cpp:141-190 for the n cameras of a frame: packed footage samples straight into the ISP (widened inside its first kernel exactly
 * as RawConverter widens them, then loadImage's 16-bit branch). isp: created with 16 output bits (Unpacker's kOutputBpp). high_byte 0:
 * out_dev[k] = h x w x 3 uint16 BGR, the image Unpacker writes as PNG; high_byte 1: h x w x 3 uint8, the renderer's input (see
 * s360_isp_process_frame_dev). Device buffers, pointer arrays on the HOST, asynchronous on the context's stream. */
int s360_isp_process_footage_dev(s360_ctx* ctx, s360_isp* isp, int n, const void* const* packed_dev, int bits, int w, int h, void* const* out_dev,
                                 int high_byte);

/* One frame through host buffers (what the Unpacker tool calls per camera and frame): packed: w * h * bits / 8 bytes; out16: h x w x 3
 * uint16 BGR; raw16_out (may be NULL): the widened samples, w x h uint16 (the --output_raw_dir dump). isp may be NULL when only the raw
 * dump is wanted (a camera without an ISP configuration, Unpacker.cpp:165-168); then out16 is not written. */
int s360_isp_process_footage(s360_ctx* ctx, s360_isp* isp, const void* packed, int bits, int w, int h, uint16_t* out16, uint16_t* raw16_out);

/* ---- whole frame ---- */

/* renderStereoPanorama (TRSP:716-972) without file I/O. images: one BGR8 image per rig
 * camera, in rig-JSON order, all img_w x img_h. equirect_out: BGR8, left eye over right eye;
 * its size is given by s360_frame_output_size. cubemap_out may be NULL. Host pointers;
 * copies in/out are part of the call. */
int s360_frame_output_size(s360_ctx* ctx, const s360_params* p, int* eq_w, int* eq_h, int* cube_w, int* cube_h);
int s360_render_frame(s360_ctx* ctx, const s360_params* p, const uint8_t* const* images_bgr, int img_w, int img_h,
                      uint8_t* equirect_out, uint8_t* cubemap_out);
/* Device-pointer variant: inputs/outputs already resident in HBM, enqueued on the context
 * stream, no synchronisation (caller uses s360_ctx_synchronize or its own events). */
int s360_render_frame_dev(s360_ctx* ctx, const s360_params* p, const uint8_t* const* images_bgr_dev, int img_w,
                          int img_h, uint8_t* equirect_out_dev, uint8_t* cubemap_out_dev);

/* ---- encode hand-off ----
 * The reference writes the equirect as PNG and lets ffmpeg convert to yuv420p and encode on the CPU (scripts/run_all.py:74-88, 313-330).
 * Here the finished BGR8 frame (s360_render_frame_dev's output, still in HBM) is converted on the device to NV12 -- a w x h luma plane
 * and an interleaved (w / 2) x (h / 2) Cb Cr plane, sizes rounded up for odd w / h -- the input surface a hardware encoder registers
 * from a device pointer. matrix: 0 = BT.601 (what swscale applies to untagged RGB), 1 = BT.709; limited range; 8-bit integer form
 * ((66 R + 129 G + 25 B + 128) >> 8) + 16 etc.; chroma from the rounded 2 x 2 mean. Pitches in bytes, 0 = tightly packed. The image has
 * no encoder library, so the hand-off ends at the NV12 surface (INTEGRATION.md shows the NVENC calls that take it). */
int s360_frame_to_nv12_dev(s360_ctx* ctx, const uint8_t* bgr_dev, size_t bgr_stride, int w, int h, int matrix, uint8_t* y_dev, size_t y_pitch,
                           uint8_t* uv_dev, size_t uv_pitch);
/* Host buffers: y_out w x h, uv_out (2 * ceil(w / 2)) x ceil(h / 2). */
int s360_frame_to_nv12(s360_ctx* ctx, const uint8_t* bgr, int w, int h, int matrix, uint8_t* y_out, uint8_t* uv_out);

/* Intermediates of the last s360_render_frame* call on this context, copied to host for
 * stage-wise parity tests. name: "projection"(idx=side cam) "flowLtoR"/"flowRtoL"(idx=pair)
 * "panoL" "panoR" "topSpherical" "bottomSpherical" "bottomMerged" "poleFlow"(0..3) "poleWarped"(0..3)
 * "eyeL" "eyeR". Query size with out == NULL. is_float: 0 = uint8, 1 = float32. */
int s360_frame_intermediate(s360_ctx* ctx, const char* name, int idx, void* out, int* w, int* h, int* channels,
                            int* is_float);
/* Per-stage device time (ms) of the last frame, measured with CUDA events on the context
 * stream, in the order of the reference's runtime breakdown (TRSP:964-971):
 * total, projection, side flow, novel view, pole flow+composite, sharpen, cubemap+resize */
int s360_frame_stage_ms(s360_ctx* ctx, float out_ms[7]);

/* ---- flow file format of saveFlowToFile/readFlowFromFile (CvUtil.cpp:159-199) ---- */
int s360_flow_save(const char* path, const float* flow, int w, int h);
int s360_flow_load(const char* path, float** flow_out, int* w, int* h); /* free with s360_free */
void s360_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* S360_H */
