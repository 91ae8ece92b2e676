/*
 * sindyn.h -- C ABI of libsindyn_cuda: B200 (sm_100a) kernels for SInDSLAM's per-frame
 * dynamic-region detection path (DynaDetect + masked ORBextractor).
 *
 * Boundary rules: extern "C", plain pointers and sizes, int return codes (0 = SINDYN_OK), no
 * exceptions, no OpenCV / torch types.  The caller owns all HOST memory; the library owns all
 * DEVICE memory.  One handle = one sequence = one CUDA device + stream; handles are independent
 * (multi-sequence / multi-GPU = several handles, no collectives).
 *
 * Every entry point cites the reference interface (file:line under qimao7213/SInDSLAM) it
 * replaces.  Image arguments are row-major with an explicit row step in BYTES (cv::Mat::step).
 */
#ifndef SINDYN_H
#define SINDYN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sindyn_ctx *sindyn_handle;

enum sindyn_status {
    SINDYN_OK = 0,
    SINDYN_ERR_INVALID = 1,   /* bad argument / size mismatch */
    SINDYN_ERR_CUDA = 2,      /* CUDA runtime error (see sindyn_last_error) */
    SINDYN_ERR_NO_DEVICE = 3, /* no CUDA device: there is NO CPU fallback */
    SINDYN_ERR_STATE = 4,     /* call order violated (e.g. detect before set_prev_frames) */
    SINDYN_ERR_CAPACITY = 5   /* fixed-capacity device list overflowed (> 128 components of one frame, > 64 planes, ...).  Reported by the
                               * call that synchronises the frame, AFTER the frame has been committed: the outputs and the inter-frame
                               * state (imgDynaLast, imgLabelLast, ...) hold the truncated result and the frame ring has rolled.  The
                               * reference has no such caps; a caller that wants to continue treats the frame as degraded, one that
                               * wants to retry restores the state with sindyn_set_state first. */
};

/* Constructor arguments of ORB_SLAM2::DynaDetect (include/DynaDetect.h:98-105) plus the
 * constants hard-coded in src/DynaDetect.cc:43-59,1029,1033 made run-time parameters.
 * sindyn_default_config fills the reference's values. */
typedef struct sindyn_config {
    int width, height;            /* DynaDetect.cc:43-44 (640x480 hard-coded there) */
    float fx, fy, cx, cy;         /* DynaDetect.h:100-103 */
    float depth_scale;            /* DynaDetect.h:104 (DepthMapFactor, raw units per metre) */
    float flow_scale;             /* DynaDetect.cc:1033  scale_element = 0.6 */
    float brox_alpha, brox_gamma, brox_pyr_scale;             /* DynaDetect.cc:1029 */
    int brox_inner, brox_outer, brox_solver;                  /* DynaDetect.cc:1029 */
    float brox_omega;             /* SOR relaxation (internal to the un-vendored NCV solver) */
    int refine;                   /* 1 = run the VariationalRefinement-equivalent pass (DynaDetect.cc:1133-1143) */
    int n_row_cluster, n_col_cluster; /* DynaDetect.cc:46  (3 x 4 = 12 clusters) */
    float depth_weight;           /* DynaDetect.cc:48 */
    int device;                   /* CUDA device ordinal */
    int use_graphs;               /* 1 = replay captured CUDA graphs for the launch-bound stages */
    int plane_edges;              /* 1 = run the PEAC plane-edge stage (DynaDetect.cc:592-593) */
    int stage_timing;             /* 1 = bracket every stage with CUDA events (sindyn_get_stage_ms) */
} sindyn_config;

void sindyn_default_config(sindyn_config *cfg, int width, int height);

/* ------------------------------------------------------------------ life cycle */
/* Replaces: DynaDetect::DynaDetect(imgLast, imgLastLast, fx, fy, cx, cy, depthScale)
 * (include/DynaDetect.h:98-126) -- construction + intrinsics. */
int sindyn_create(const sindyn_config *cfg, sindyn_handle *out);
int sindyn_destroy(sindyn_handle h);
const char *sindyn_last_error(sindyn_handle h);
/* Use an external CUDA stream (cudaStream_t) for all work of this handle, e.g. torch's current
 * stream so that torch.cuda.Event timing brackets the kernels.  NULL restores the own stream. */
int sindyn_set_stream(sindyn_handle h, void *cuda_stream);
int sindyn_synchronize(sindyn_handle h);
/* number of kernel launches issued by this handle since creation (graph nodes count individually) */
unsigned long long sindyn_launch_count(sindyn_handle h);

/* The two BGR frames the constructor receives (DynaDetect.h:98-105; driver: rgbd_tum_noros.cc:103-107).
 * Also zeroes the inter-frame state (imgDynaLast, imgMaskHighErrorLast, imgLabelLast; DynaDetect.h:109-112). */
int sindyn_set_prev_frames(sindyn_handle h, const uint8_t *bgr_last, size_t step_last,
                           const uint8_t *bgr_lastlast, size_t step_lastlast);

/* Replaces: DynaDetect::DetectDynaArea(img, imgDepth, imgDyna, imgLabel, nImg)
 * (include/DynaDetect.h:127-131, src/DynaDetect.cc:1377-1666).
 * bgr: 8UC3, depth: 16UC1 raw; mask_out: 8UC1 {0,125,255}; label_out: 8UC1 merged cluster ids.
 * Host pointers; H2D/D2H copies and the final synchronize happen inside. */
int sindyn_detect(sindyn_handle h, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step,
                  uint8_t *mask_out, size_t mask_step, uint8_t *label_out, size_t label_step, int frame_idx);

/* Device-resident variant used for kernel-only timing: frames are staged once into `slot`
 * (0 <= slot < SINDYN_MAX_SLOTS) and detect runs without host traffic; results stay on the
 * device until sindyn_get_buffer. */
#define SINDYN_MAX_SLOTS 512
int sindyn_upload_frame(sindyn_handle h, int slot, const uint8_t *bgr, size_t bgr_step,
                        const uint16_t *depth, size_t depth_step);
int sindyn_detect_resident(sindyn_handle h, int slot, int frame_idx);
/* Copy out the device-resident results of the last detect call (either pointer may be NULL): mask / labels W x H u8.
 * Synchronizes, checks the fixed-capacity lists and collects the per-stage timings. */
int sindyn_get_detect_results(sindyn_handle h, uint8_t *mask, uint8_t *labels);

/* Replaces: DynaDetect::DetectDynaByDenseOpticalFLow(std::promise<stImgMasks>&)
 * (include/DynaDetect.h:147, src/DynaDetect.cc:1023-1374): the flow branch the reference runs in its
 * own std::thread -- gray/resize, Brox, large-motion test, refinement, up-sampling, sample weighting,
 * homography, residual, Otsu/Triangle thresholds.  Returns stImgMasks (DynaDetect.h:16-30):
 * mask_low (0/128) and mask_high (0/255), W x H u8 host buffers (either may be NULL).
 * roll != 0 additionally rolls imgRGBLastLast <- imgRGBLast <- cur (DynaDetect.cc:1661-1662) so that a
 * sequence can be streamed through this entry point alone (BASELINE config "flow + residual"). */
int sindyn_flow_residual(sindyn_handle h, const uint8_t *bgr, size_t bgr_step, uint8_t *mask_low, uint8_t *mask_high, int roll);
/* Same on a frame staged with sindyn_upload_frame: no host traffic and no synchronisation -- with use_graphs the
 * whole flow branch is one CUDA graph whose large-motion decision (taken on the host by the reference,
 * DynaDetect.cc:1073-1114) is a conditional node; the flag is reported by sindyn_get_flow_results. */
int sindyn_flow_residual_resident(sindyn_handle h, int slot, int roll);
/* Copy out the device-resident results of the last flow_residual call (any pointer may be NULL):
 * flow W x H x 2 float, H 3x3 double, thresholds[4], masks. */
int sindyn_get_flow_results(sindyn_handle h, float *flow, double *H_out, float *thresholds, uint8_t *mask_low, uint8_t *mask_high,
                            int *large_motion);
/* Diagnostics (no reference equivalent): which launch path the LAST whole-frame call (sindyn_detect / sindyn_flow_residual*)
 * took.  info[0] = 1 if the flow branch ran as the one captured CUDA graph whose large-motion decision
 * (DynaDetect.cc:1097-1114) is a conditional node, 0 = classic path with the host decision; info[1] = 1 if the one-graph
 * path has been disabled for this handle (capture failed); info[2] = 1 if the clustering branch was replayed as a graph;
 * info[3] = reserved.  Used by the tests that compare the graph path with the classic path bit for bit. */
int sindyn_get_path_info(sindyn_handle h, int info[4]);

/* Measurement hook (no reference equivalent): runs one Brox solve on the handle's resident frames
 * WITHOUT the CUDA graph, bracketing every launch of the tiled SOR kernel (k_brox_sor) with CUDA events
 * on the handle's stream.  out[0] = summed SOR-kernel ms, out[1] = number of SOR launches,
 * out[2] = whole-solve ms, out[3] = pixel-sweeps those launches processed (sum over launches of
 * w*h*sweeps of the launch). */
int sindyn_brox_profile(sindyn_handle h, double *out4);

/* Driver post-step: 15x15 ellipse dilation of the mask (rgbd_tum_noros.cc:108,136-139), and the
 * generic getStructuringElement(MORPH_ELLIPSE,k x k) morphology used throughout DynaDetect.cc:51-59.
 * op: 0 = dilate, 1 = erode, 2 = open, 3 = close. k in {1..15}. */
int sindyn_morph_ellipse(sindyn_handle h, const uint8_t *src, size_t src_step, uint8_t *dst, size_t dst_step,
                         int width, int height, int k, int op);

/* ------------------------------------------------------------------ stage-level entry points
 * (the seams of SURVEY.md section 4: identical inputs can be injected at every stage boundary,
 * like the authors' own .flo injection hook, DynaDetect.cc:1149-1158). */

/* cv::cuda::BroxOpticalFlow::calc(I0 = current, I1 = older, flow) (DynaDetect.cc:1029,1072,1124).
 * I0/I1: w x h float in [0,1], densely packed; flow_uv: w x h x 2 float (CV_32FC2 layout), the RAW
 * solver sign (I0(x) ~ I1(x + w)). */
int sindyn_flow_brox(sindyn_handle h, const float *I0, const float *I1, int w, int hgt, float *flow_uv);

/* Flow-branch prologue (DynaDetect.cc:1390-1392,1033-1048): BGR->gray, resize to
 * (int)(flow_scale*W) x (int)(flow_scale*H) INTER_LINEAR, u8 out (and /255 float inside). */
int sindyn_gray_resize(sindyn_handle h, const uint8_t *bgr, size_t bgr_step, uint8_t *gray_full, uint8_t *gray_small);

/* Full flow branch up to the up-sampled flow (DynaDetect.cc:1033-1147): gray/resize of the three
 * frames, Brox(cur, lastlast), large-motion test, optional Brox(cur, last), negate, refinement,
 * resize to W x H, scale by 1/flow_scale.  flow_out: W x H x 2 float. large_motion_out may be NULL. */
int sindyn_flow_branch(sindyn_handle h, const uint8_t *bgr_cur, size_t step_cur, float *flow_out, int *large_motion_out);

/* cv::VariationalRefinement::create()->calc(I0, I1, flow) with default parameters
 * (DynaDetect.cc:1133-1143): I0/I1 u8 w x h, flow in/out w x h x 2 float. */
int sindyn_flow_refine(sindyn_handle h, const uint8_t *I0, const uint8_t *I1, int w, int hgt, float *flow_uv);

/* Replaces: cv::findHomography(srcPoints, dstPoints, cv::noArray(), cv::RHO) (the call of DynaDetect.cc:1235; OpenCV calib3d
 * rho.cpp, un-vendored) on an arbitrary ORDERED list of n <= 4096 correspondences (x, y floats; the PROSAC sampler of RHO
 * consumes the order).  H_out: 3x3 row-major doubles, bit-identical to the library's result for n >= 5 (all zeros when fewer
 * than 4 inliers were found); inlier_mask_out (optional): n bytes; info_out (optional, 12 ints): [0] n, [1] inliers of the best
 * model, [2] models evaluated, [3] refinement iterations, [4..6] diagnostics: kilo-cycles of the sampling loop, of the
 * non-randomness optimisation and of the refinement, [7] PROSAC iterations, [8..10] kilo-cycles of the control thread in
 * sampling + 4-point solve / reprojection tests + SPRT scan / bookkeeping, [11] models that needed the sequential SPRT fallback. */
int sindyn_find_homography_rho(sindyn_handle h, const float *src_xy, const float *dst_xy, int n, double *H_out, uint8_t *inlier_mask_out,
                               int *info_out);
/* Sample weighting + sort + in-border filter (DynaDetect.cc:1163-1231) followed by the robust
 * homography that replaces cv::findHomography(pts, ptsLast, noArray(), RHO) (DynaDetect.cc:1235).
 * flow: W x H x 2 float (already up-sampled).  H_out: 3x3 row-major double.  Uses the handle's
 * state images (imgDynaLast / imgLabelLast).  n_pairs_out may be NULL. */
int sindyn_estimate_homography(sindyn_handle h, const float *flow, double *H_out, int *n_pairs_out);
/* Only the sample list (for parity of DynaDetect.cc:1163-1231): pts / pts_last are n x 2 float. */
int sindyn_sample_pairs(sindyn_handle h, const float *flow, float *pts, float *pts_last, int capacity, int *n_out);

/* Residual against the homography flow + thresholds -> two masks
 * (DynaDetect.cc:1236-1367): residual_mag_out (W x H float, may be NULL), mask_low (0/128),
 * mask_high (0/255); thresholds_out[4] = {otsu, triangle, t_low, t_high} (may be NULL). */
int sindyn_residual_homography(sindyn_handle h, const float *flow, const double *Hm, float *residual_mag_out,
                               uint8_t *mask_low, uint8_t *mask_high, float *thresholds_out);
/* north_star variant: depth back-projection with the SE(3) pose T_old_cur (3x4 row-major, maps
 * current-camera points into the older camera), reprojection to predicted flow, residual against
 * the measured flow; same thresholding.  Needs an interface extension upstream (pose in). */
int sindyn_residual_pose(sindyn_handle h, const float *flow, const uint16_t *depth, size_t depth_step,
                         const double *T_old_cur, float *residual_mag_out, uint8_t *mask_low, uint8_t *mask_high,
                         float *thresholds_out);

/* DynaDetect::SegByKmeans (DynaDetect.cc:315-420): 4-level pyramid K-means on (X,Y,1.5 Z).
 * labels_out: W x H u8; points_out: N x 3 float (may be NULL); centers_out: 12 x 3 float (may be NULL).
 * Uses the handle's imgLabelLast for initialisation. */
int sindyn_kmeans(sindyn_handle h, const uint16_t *depth, size_t depth_step, uint8_t *labels_out, float *points_out,
                  float *centers_out);

/* Cluster ordering (DynaDetect.cc:1425-1491): runs on the last sindyn_kmeans result.
 * label_for_seg_edge_out: W x H u8 (0/255, dilated 7x7); order_out[12] = cluster ids kept, depth
 * order, -1 padded; n_kept_out = allLabels.size(). */
int sindyn_cluster_order(sindyn_handle h, uint8_t *label_for_seg_edge_out, int *order_out, int *n_kept_out);

/* DynaDetect::CalOccluded gradient edges + end points (DynaDetect.cc:434-536).
 * total_area_out (0/255), grad_edges_out (0/255, after OPEN 4x4); endpoints_out: n x 2 int (x,y)
 * after NMS, raster order. */
int sindyn_depth_edges(sindyn_handle h, const uint16_t *depth, size_t depth_step, uint8_t *total_area_out,
                       uint8_t *grad_edges_out, int *endpoints_out, int capacity, int *n_endpoints_out);

/* PEAC plane-contour edges (DynaDetect.cc:558-593; include/PEAC plane_fitter_pcl.hpp:275-317).
 * plane_edges_out: W x H u8 (imgEdgeByPlane before filtering). */
int sindyn_plane_edges(sindyn_handle h, const uint16_t *depth, size_t depth_step, uint8_t *plane_edges_out);

/* Test hook: final plane membership image of the last PEAC run (W x H ints: final plane id or -1), the planes extracted
 * by the block-level clustering (n x 3 ints: root block id, point count, final plane id; capacity 64), and the counts. */
int sindyn_get_peac_debug(sindyn_handle h, int *membership_out, int *planes_rid_n_final, int *n_planes_out, int *n_final_out);

/* Plane-edge filtering + imgOccluded1/2 (DynaDetect.cc:598-641). Inputs: plane edges (raw),
 * gradient edges, end points. Outputs occluded1 (all edges, CLOSE 3x3) and occluded2 (plane edges kept). */
int sindyn_filter_plane_edges(sindyn_handle h, const uint8_t *plane_edges, const uint8_t *grad_edges,
                              const int *endpoints, int n_endpoints, uint8_t *occluded1_out, uint8_t *occluded2_out);

/* DynaDetect::SegAndMergeV2 (DynaDetect.cc:653-1018) on the last k-means / cluster-order result.
 * occluded1/2 and label_for_seg_edge are injected (host); label_out: W x H u8 merged labels
 * (0 = invalid, 1..n). */
int sindyn_recluster(sindyn_handle h, const uint8_t *occluded1, const uint8_t *occluded2, const uint16_t *depth,
                     size_t depth_step, uint8_t *label_out, int *n_components_out);

/* Test hook (no reference equivalent): RAG matrix of the last sindyn_recluster / sindyn_detect call
 * (correlationMatrixTotal, DynaDetect.cc:894; (n+1)^2 floats in sorted-rank space) and per-component area / score /
 * sorted order.  Returns the number of components n (>= 0) or a negative... status is never negative: returns n. */
int sindyn_get_recluster_debug(sindyn_handle h, float *T_out, int t_capacity, int *area_out, float *score_out, int *order_out);

/* Mask fusion + per-cluster decision + final mask (DynaDetect.cc:1553-1636) and state roll
 * (DynaDetect.cc:1660-1664).  Inputs injected: low (0/128), high (0/255), total_area, labels. */
int sindyn_dynamic_decide(sindyn_handle h, const uint8_t *mask_low, const uint8_t *mask_high, const uint8_t *total_area,
                          const uint8_t *labels, uint8_t *dyna_out);

/* Inter-frame state (DynaDetect.h:164-178): which = 0 imgDynaLast, 1 imgMaskHighErrorLast,
 * 2 imgLabelLast (u8 W x H); 3 imgRGBLast, 4 imgRGBLastLast (u8 W x H x 3). */
int sindyn_get_state(sindyn_handle h, int which, uint8_t *out);
int sindyn_set_state(sindyn_handle h, int which, const uint8_t *in);

/* Per-stage device milliseconds of the last sindyn_detect (CUDA events):
 * [0] upload+gray/resize [1] brox [2] large-motion+refine+upsample [3] homography [4] residual+masks
 * [5] kmeans [6] depth edges [7] plane edges [8] recluster [9] decide+final [10] total.  n <= 16. */
int sindyn_get_stage_ms(sindyn_handle h, float *ms, int n);

/* ------------------------------------------------------------------ ORB extractor */
typedef struct sindyn_orb *sindyn_orb_handle;
typedef struct sindyn_keypoint { /* fields of cv::KeyPoint used by ORB-SLAM2 */
    float x, y, size, angle, response;
    int octave;
} sindyn_keypoint;

/* Replaces ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (include/ORBextractor.h:54-55, src/ORBextractor.cc:410-470). */
int sindyn_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast,
                      int width, int height, int device, sindyn_orb_handle *out);
int sindyn_orb_destroy(sindyn_orb_handle h);
/* Replaces ORBextractor::operator()(image, mask, keypoints, descriptors)
 * (include/ORBextractor.h:62-64, src/ORBextractor.cc:1043-1164). mask may be NULL (== empty Mat).
 * kps: capacity entries; desc: capacity x 32 bytes. *n_out = number of keypoints written. */
int sindyn_orb_extract(sindyn_orb_handle h, const uint8_t *gray, size_t gray_step, const uint8_t *mask, size_t mask_step,
                       sindyn_keypoint *kps, uint8_t *desc, int capacity, int *n_out);
/* One iteration of the driver loop, fused: replaces rgbd_tum_noros.cc:132-139 (DynaDetect::DetectDynaArea followed by the
 * dilate_k x dilate_k elliptic dilation of the mask, 15 in the reference) plus the part of System::TrackRGBD that produces the
 * key points: Tracking::GrabImageRGBD's colour conversion (Tracking.cc:246-252; rgb_order != 0 = Camera.RGB: 1 = CV_RGB2GRAY,
 * 0 = CV_BGR2GRAY) and Frame::ExtractORB2 -> ORBextractor::operator()(gray, dilated mask) (Frame.cc:300-317,
 * ORBextractor.cc:1043-1164).  Same results as sindyn_detect + sindyn_morph_ellipse + sindyn_orb_extract called one after the
 * other; the frame is uploaded once, the mask never leaves the device between the stages, and the mask-independent half of
 * the extraction overlaps the detection.  mask_out receives the DILATED mask (what the driver hands to TrackRGBD). */
int sindyn_track_frame(sindyn_handle h, sindyn_orb_handle o, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step,
                       int rgb_order, int dilate_k, uint8_t *mask_out, size_t mask_step, uint8_t *label_out, size_t label_step,
                       sindyn_keypoint *kps, uint8_t *desc, int capacity, int *n_out, int frame_idx);
/* Same on a frame staged with sindyn_upload_frame: no host traffic, no synchronisation (kernel-only timing; the extractor's
 * stream is joined into the detector handle's stream).  Results via sindyn_track_get_results. */
int sindyn_track_frame_resident(sindyn_handle h, sindyn_orb_handle o, int slot, int rgb_order, int dilate_k, int frame_idx);
/* sindyn_track_frame_resident only enqueues (on several streams: with CUDA graphs on, the image-only stages of frame i + 1 --
 * gray / resize, Brox flow, refinement, PEAC plane fitter, the unmasked half of the extractor -- overlap the decision of frame
 * i, the software pipeline of pipe.cu; the reference's driver loop, rgbd_tum_noros.cc:113-192, hands over one frame after the
 * other and nothing in those stages reads the detector's state).  After sindyn_track_join everything enqueued so far precedes
 * the next operation on the detector handle's stream, and the next frame follows whatever that stream holds by then. */
int sindyn_track_join(sindyn_handle h, sindyn_orb_handle o);
/* Asynchronous form of sindyn_track_frame for a caller that has the next image at hand while it still works on the current one
 * (rgbd_tum_noros.cc:113-192 reads a recorded sequence; Tracking::GrabImageRGBD, Tracking.cc:209-240, needs mask + key points of
 * frame i only).  sindyn_track_submit uploads frame i + 1 and enqueues all of its work without waiting; sindyn_track_collect
 * returns the oldest submitted frame (same outputs as sindyn_track_frame; mask_out / label_out / kps / desc may be NULL).  At
 * most SINDYN_TRACK_MAX_IN_FLIGHT (3) frames may be in flight (SINDYN_ERR_STATE otherwise); results are bit-identical to sindyn_track_frame.  bgr / depth
 * must stay valid until the frame is collected if they are pinned (page-locked) memory, pageable memory is copied at once. */
#define SINDYN_TRACK_MAX_IN_FLIGHT 3
int sindyn_track_submit(sindyn_handle h, sindyn_orb_handle o, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step,
                        int rgb_order, int dilate_k, int frame_idx);
int sindyn_track_collect(sindyn_handle h, sindyn_orb_handle o, uint8_t *mask_out, size_t mask_step, uint8_t *label_out, size_t label_step,
                         sindyn_keypoint *kps, uint8_t *desc, int capacity, int *n_out);
int sindyn_track_get_results(sindyn_handle h, sindyn_orb_handle o, uint8_t *mask_dilated, uint8_t *labels, sindyn_keypoint *kps, uint8_t *desc,
                             int capacity, int *n_out);

/* mvImagePyramid[level] (include/ORBextractor.h:88): copies level image (without the 19-px pad). */
int sindyn_orb_get_pyramid_level(sindyn_orb_handle h, int level, uint8_t *out, int *w_out, int *h_out);
/* Test hook: FAST candidates of one level after the last extract, in distribution order (vToDistributeKeys,
 * ORBextractor.cc:820-825): xyr = n x 3 ints (x, y relative to minBorder, response). */
int sindyn_orb_get_candidates(sindyn_orb_handle h, int level, int *xyr, int capacity, int *n_out);
/* Test hook: raw device planes of one level: which = 0 padded image (w+38 x h+38), 1 FAST score map, 2 blurred image. */
int sindyn_orb_get_plane(sindyn_orb_handle h, int level, int which, uint8_t *out, int *w_out, int *h_out);
int sindyn_orb_set_stream(sindyn_orb_handle h, void *cuda_stream);
unsigned long long sindyn_orb_launch_count(sindyn_orb_handle h);
const char *sindyn_orb_last_error(sindyn_orb_handle h);

/* SURVEY.md 8(f) row f2 -- what ORB_SLAM2::Frame does with the keypoints right after the extraction (src/Frame.cc:143-170),
 * on the keypoints that are still resident on the device: UndistortKeyPoints (Frame.cc:714-753), ComputeImageBounds
 * (:507-535), ComputeStereoFromRGBD (depth lookup at the distorted keypoint, virtual right coordinate u - bf/d) and
 * AssignFeaturesToGrid / PosInGrid (:453-463; 64 x 48 cells).
 * depth_raw: 16UC1 raw depth of the frame; depth_map_factor = 1 / DepthMapFactor (Tracking.cc:142-146).
 * Outputs (any may be NULL): keys_un n x 2 (mvKeysUn), depth n (mvDepth, -1 = none), u_right n (mvuRight), bounds[4] =
 * {mnMinX, mnMaxX, mnMinY, mnMaxY}, grid as CSR: cell (i, j) = mGrid[i][j] holds grid_indices[grid_offsets[i*48+j] ..
 * grid_offsets[i*48+j+1]) in push_back order. */
typedef struct sindyn_frame_params {
    float fx, fy, cx, cy;          /* mK */
    float k1, k2, p1, p2, k3;      /* mDistCoef (Tracking.cc:60-75) */
    float bf;                      /* mbf */
    float depth_map_factor;        /* mDepthMapFactor */
} sindyn_frame_params;
int sindyn_orb_frame_features(sindyn_orb_handle h, const uint16_t *depth_raw, size_t depth_step, const sindyn_frame_params *params,
                              float *keys_un, float *depth_out, float *u_right_out, float *bounds_out, int *grid_offsets,
                              int *grid_indices, int capacity, int *n_out);

/* ---- "next" row f4: frame-to-frame descriptor matching of Tracking::TrackWithMotionModel.
 * Replaces ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
 * (src/ORBmatcher.cc:1328-1470) with ORBmatcher::DescriptorDistance (:1647-1665), ComputeThreeMaxima (:1601-1643) and
 * Frame::GetFeaturesInArea (src/Frame.cc:398-452).
 * CurrentFrame = the frame resident in the handle (last sindyn_orb_extract + sindyn_orb_frame_features: mvKeysUn, mvuRight,
 * mDescriptors, mGrid, image bounds, mvScaleFactors).  LastFrame arrives as arrays over its n_last key points:
 *   last_valid[i]    = (mvpMapPoints[i] != NULL && !mvbOutlier[i])
 *   last_xyz_world   = pMP->GetWorldPos()        (n_last x 3 float)
 *   last_desc        = pMP->GetDescriptor()      (n_last x 32 bytes)
 *   last_octave[i]   = LastFrame.mvKeys[i].octave,  last_angle[i] = LastFrame.mvKeysUn[i].angle
 *   last_observed[i] = (pMP->Observations() > 0)
 * cur_blocked (optional, n_cur bytes) = (CurrentFrame.mvpMapPoints[i2] != NULL && Observations() > 0) on entry; NULL = none
 * (TrackWithMotionModel clears the vector first, Tracking.cc:876).
 * match_out[i2] = index i of the last-frame point whose map point was assigned to current key point i2, or -1;
 * *nmatches_out = the function's return value.  Poses are row-major 4 x 4 floats (cv::Mat CV_32F mTcw). */
typedef struct sindyn_match_params {
    float fx, fy, cx, cy;          /* CurrentFrame intrinsics */
    float bf, b;                   /* mbf, mb */
    float Tcw_cur[16], Tcw_last[16];
    float th;                      /* search radius factor (15 for RGB-D, Tracking.cc:884) */
    int mono;                      /* bMono */
    int check_orientation;         /* ORBmatcher::mbCheckOrientation */
} sindyn_match_params;
int sindyn_orb_search_by_projection(sindyn_orb_handle h, const sindyn_match_params *params, int n_last, const float *last_xyz_world,
                                    const uint8_t *last_valid, const uint8_t *last_desc, const int *last_octave, const float *last_angle,
                                    const uint8_t *last_observed, const uint8_t *cur_blocked, int *match_out, int capacity,
                                    int *n_cur_out, int *nmatches_out);

/* ---- "next" row f3: per-keyframe point-cloud generation of the dense-map consumer
 * (octomap_pub/src/pubPointCloud.cc, SubscribeAndPublish::generatePointCloud).  The octree insertion itself
 * (octomap::ColorOcTree) is third-party host code and stays with the caller.
 * Point layout: 16 bytes (x, y, z in the world frame, b, g, r); NaN coordinates mark rejected pixels exactly as the
 * reference's non-dense pcl::PointCloud does.  Poses are row-major 4 x 4 doubles.  intr5 = {fx, fy, cx, cy, depthScale}
 * as the node reads them (doubles); NULL = the handle's configuration. */
typedef struct sindyn_point {
    float x, y, z;
    uint8_t b, g, r, a;
} sindyn_point;
/* Single-frame overload (pubPointCloud.cc:392-470): every 3rd pixel; mask >= 240 or depth outside [0.01, 10] m -> NaN
 * point; colour from the BGR image; pcl::transformPointCloud by Twc.  points_out holds ceil(H/3)*ceil(W/3) points. */
int sindyn_cloud_single(sindyn_handle h, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step,
                        const uint8_t *mask, size_t mask_step, const double *Twc16, const double *intr5, sindyn_point *points_out,
                        int *n_out);
/* Cross-frame consistency overload (pubPointCloud.cc:471-678): every 2nd pixel is re-projected into the previous key frame
 * (T_rel = poseRelative), a pixel votes "occluded" for its K-means cluster when the depths disagree by more than 13 % or the
 * previous mask was dynamic (> 240); clusters i >= 1 with occlusion_i * 9 > 0.4 * |label == i| are dropped from the cloud
 * and painted 255 into mask_new (imgDynaMaskNew); the kept clusters are concatenated in cluster order, raster order inside
 * a cluster, and transformed by Twc.  points_out holds up to ceil(H/2)*ceil(W/2) points.
 * stats36_out (optional) = occlusion[12] | countNonZero(label == i)[12] | kept[12]; depth_new_out (optional) = imgDepthNew. */
int sindyn_cloud_consistent(sindyn_handle h, const uint8_t *bgr, size_t bgr_step, const uint16_t *depth, size_t depth_step,
                            const uint16_t *depth_last, size_t depth_last_step, const uint8_t *mask, size_t mask_step,
                            const uint8_t *mask_last, size_t mask_last_step, const uint8_t *label, size_t label_step,
                            const double *T_rel16, const double *Twc16, const double *intr5, sindyn_point *points_out, int *n_out,
                            uint8_t *mask_new_out, size_t mask_new_step, int *stats36_out, uint16_t *depth_new_out);

const char *sindyn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SINDYN_H */
