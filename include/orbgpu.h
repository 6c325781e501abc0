/* orbgpu.h — C ABI of the B200-native ORB front-end (liborbgpu.so).
 *
 * This is the drop-in boundary for the two hot-path classes of ORB-SLAM2:
 *   ORB_SLAM2::ORBextractor   (/root/reference/include/ORBextractor.h:45-111, src/ORBextractor.cc)
 *   ORB_SLAM2::ORBmatcher     (/root/reference/include/ORBmatcher.h:37-102,  src/ORBmatcher.cc)
 * The C++ shells in orb_slam2_with_comment_b200/csrc/host/ keep those class signatures and forward to these entry
 * points; INTEGRATION.md shows the binding a maintainer of the reference adds.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a HOST pointer unless the name ends in _dev;
 *   - every function returns 0 on success and a negative orbgpu_status otherwise;
 *     orbgpu_last_error() returns a thread-local description of the last failure;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with ORBGPU_ERR_CUDA;
 *   - handles are not thread-safe, distinct handles may be used concurrently from distinct threads
 *     (the reference runs the left and right extractor on two threads, Frame.cc:78-81).
 */
#ifndef ORBGPU_H
#define ORBGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum orbgpu_status {
    ORBGPU_OK = 0,
    ORBGPU_ERR_ARG = -1,      /* bad argument / unsupported geometry           */
    ORBGPU_ERR_CUDA = -2,     /* CUDA runtime error or no device                */
    ORBGPU_ERR_CAPACITY = -3  /* a caller-provided or configured capacity is too small */
} orbgpu_status;

/* Same 28-byte layout as cv::KeyPoint (pt.x, pt.y, size, angle, response, octave, class_id), so a
 * std::vector<cv::KeyPoint> can be filled with one memcpy (ORBextractor.cc:1072-1103). */
typedef struct orbgpu_keypoint {
    float x, y;       /* level-0 image coordinates                                  (:1095-1101) */
    float size;       /* (int)(31 * mvScaleFactor[octave])                          (:837)       */
    float angle;      /* IC_Angle, degrees in [0,360)                               (:77-104)    */
    float response;   /* FAST score                                                 (:809-815)   */
    int32_t octave;   /* pyramid level                                              (:845)       */
    int32_t class_id; /* always -1                                                               */
} orbgpu_keypoint;

const char* orbgpu_last_error(void);
int orbgpu_device_count(int* count);
/* ABI version of this header; bumped when a signature changes. */
int orbgpu_abi_version(void);
/* Process-wide default device of the C++ shells: ORBextractor::SetDevice stores it here, and the extractor, matcher and
 * vocabulary shells all create their handles on it — so extraction, BoW and matching of one process share a GPU and the
 * device-side chaining (orbgpu_frame_set_from_extraction) never meets handles of different devices.  Default 0. */
int orbgpu_set_default_device(int device);
int orbgpu_default_device(void);

/* ------------------------------------------------------------------------------------------------
 * Extraction — replaces ORBextractor (ORBextractor.h:45-111)
 * ---------------------------------------------------------------------------------------------- */
typedef struct orbgpu_extractor orbgpu_extractor;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) (ORBextractor.cc:410-470)
 * plus the device placement and the workspace bounds a GPU implementation needs up front:
 * images up to max_width x max_height, up to max_batch frames per call. */
int orbgpu_extractor_create(orbgpu_extractor** out, int device, int nfeatures, float scale_factor, int nlevels,
                            int ini_th_fast, int min_th_fast, int max_width, int max_height, int max_batch);
int orbgpu_extractor_destroy(orbgpu_extractor* ex);

/* Scale tables as the reference getters return them (ORBextractor.h:63-83): scales[0..L) = mvScaleFactor,
 * [L..2L) = mvInvScaleFactor, [2L..3L) = mvLevelSigma2, [3L..4L) = mvInvLevelSigma2; features_per_level[L]
 * = mnFeaturesPerLevel; umax[16].  Any pointer may be NULL. */
int orbgpu_extractor_tables(const orbgpu_extractor* ex, float* scales, int32_t* features_per_level, int32_t* umax);

/* The same tables without a handle or a device (pure host arithmetic of the constructor, ORBextractor.cc:413-469):
 * the C++ shell needs them at construction, before the first image fixes the workspace size. */
int orbgpu_extractor_static_tables(int nfeatures, float scale_factor, int nlevels, float* scales, int32_t* features_per_level,
                                   int32_t* umax);

/* Upper bound on the keypoints one frame can yield (nfeatures + 3*nlevels): size kp/desc buffers with it. */
int orbgpu_extractor_max_keypoints(const orbgpu_extractor* ex);

/* ORBextractor::operator() (ORBextractor.cc:1043-1105) for one 8-bit single-channel image.
 * kp_out[kp_capacity], desc_out[kp_capacity*32]; *n_out receives the keypoint count.
 * An empty image (NULL or 0 x 0) yields *n_out = 0 and success, like the reference's early return (:1046). */
int orbgpu_extract(orbgpu_extractor* ex, const uint8_t* image, int width, int height, size_t row_stride,
                   orbgpu_keypoint* kp_out, uint8_t* desc_out, int kp_capacity, int* n_out);

/* The same for `batch` independent frames of equal size (frame f starts at images + f*frame_stride).
 * Frame f's results land at kp_out + f*kp_capacity and desc_out + f*kp_capacity*32; counts[f] = its count. */
int orbgpu_extract_batch(orbgpu_extractor* ex, const uint8_t* images, int batch, int width, int height,
                         size_t row_stride, size_t frame_stride, orbgpu_keypoint* kp_out, uint8_t* desc_out,
                         int kp_capacity, int32_t* counts);

/* Colour input: what Tracking::GrabImageMonocular / GrabImageStereo / GrabImageRGBD do before the Frame constructor
 * (Tracking.cc:173-198, :214-228) fused in front of the extraction — cvtColor(im, mImGray, CV_RGB2GRAY | CV_BGR2GRAY |
 * CV_RGBA2GRAY | CV_BGRA2GRAY) with OpenCV's 8-bit fixed point (Y = (R*9798 + G*19235 + B*3735 + 2^14) >> 15, bit-exact with
 * OpenCV 4.13).  `images` holds `channels` (3 or 4) interleaved bytes per pixel, row_stride / frame_stride in bytes;
 * rgb_order != 0 means channel 0 is red (mbRGB), else blue.  gray_out (may be NULL) receives mImGray, width x height bytes
 * per frame, packed.  This entry is not pipelined; the throughput path takes gray images. */
int orbgpu_extract_batch_color(orbgpu_extractor* ex, const uint8_t* images, int batch, int width, int height, int channels,
                               int rgb_order, size_t row_stride, size_t frame_stride, uint8_t* gray_out, orbgpu_keypoint* kp_out,
                               uint8_t* desc_out, int kp_capacity, int32_t* counts);

/* Device-resident variant: all pointers are device pointers on the extractor's device, the work is enqueued on
 * the extractor's stream and the call returns without synchronising (orbgpu_extractor_sync waits). */
int orbgpu_extract_batch_dev(orbgpu_extractor* ex, const uint8_t* images_dev, int batch, int width, int height,
                             size_t row_stride, size_t frame_stride, orbgpu_keypoint* kp_out_dev,
                             uint8_t* desc_out_dev, int kp_capacity, int32_t* counts_dev);
int orbgpu_extractor_sync(orbgpu_extractor* ex);
/* cudaStream_t of the extractor (as void*), so callers can record events on the launching stream. */
int orbgpu_extractor_stream(orbgpu_extractor* ex, void** stream_out);
/* Number of kernels the last extract call launched (for bench.py's gpu_launches). */
int orbgpu_extractor_last_launches(const orbgpu_extractor* ex);

/* Per-stage device time of the last extract call, measured with CUDA events on the extractor's stream:
 * ms5 = { pyramid (level 0 + 7 resizes), FAST cells, octree, blur (8 levels), orientation+descriptors }. */
int orbgpu_extractor_set_profiling(orbgpu_extractor* ex, int enable);
/* The 19-px BORDER_REFLECT_101 frame around every pyramid level (copyMakeBorder, ORBextractor.cc:1122-1128).  Nothing in
 * operator() — or anywhere else in ORB-SLAM2 — reads it: FAST cells start 16 px inside a level, the key-point windows stay inside
 * it, GaussianBlur runs on an isolated clone (:1085-1086).  By default it is therefore written when a bordered level is read
 * back (orbgpu_extractor_read_level with bordered != 0); enable != 0 writes it with every extraction call, as the reference
 * does (about 5 % of the extraction time).  The bytes are the same either way. */
int orbgpu_extractor_set_eager_frame(orbgpu_extractor* ex, int enable);
int orbgpu_extractor_stage_ms(orbgpu_extractor* ex, float* ms5);

/* mvImagePyramid (ORBextractor.h:86, filled by ComputePyramid :1107-1132) of frame `frame` of the last call.
 * bordered != 0 copies the (w+38) x (h+38) buffer with its 19-px BORDER_REFLECT_101 frame, else the w x h level. */
int orbgpu_extractor_level_dims(const orbgpu_extractor* ex, int level, int* width, int* height);
int orbgpu_extractor_read_level(orbgpu_extractor* ex, int frame, int level, int bordered, uint8_t* out,
                                size_t out_stride);

/* Stage taps of the last call, for stage-level parity tests:
 *   stage 0: FAST candidates of `level` in emission order, coordinates relative to (minBorderX,minBorderY)
 *            (= vToDistributeKeys, ORBextractor.cc:778-826); only x, y, response are meaningful
 *   stage 1: keypoints of `level` after DistributeOctTree + orientation, level coordinates (:831-852)
 *   stage 2 is read with orbgpu_extractor_read_blurred (the GaussianBlur'ed level, :1085-1086). */
int orbgpu_extractor_read_points(orbgpu_extractor* ex, int frame, int level, int stage, orbgpu_keypoint* out,
                                 int capacity, int* n_out);
int orbgpu_extractor_read_blurred(orbgpu_extractor* ex, int frame, int level, uint8_t* out, size_t out_stride);

/* Stand-alone DistributeOctTree (ORBextractor.cc:539-763) on the device, for stage-level parity tests:
 * n candidates (x, y, response used), rectangle [min_x,max_x) x [min_y,max_y), quota n_features. */
int orbgpu_octree(orbgpu_extractor* ex, const orbgpu_keypoint* candidates, int n, int min_x, int max_x, int min_y,
                  int max_y, int n_features, orbgpu_keypoint* out, int capacity, int* n_out);
/* Which of the two device formulations of DistributeOctTree answered the last orbgpu_octree call: 1 = the pass-free
 * construction (cell histogram, csrc/og_octree2.cuh), 0 = the division-pass state machine it hands deep trees to
 * (csrc/og_octree.cuh).  Both reproduce ORBextractor.cc:539-763 exactly; tests use this to cover each of them. */
int orbgpu_octree_last_path(const orbgpu_extractor* ex);

/* Frame::ComputeStereoMatches (Frame.cc:501-675) for every stereo pair of the last call: `left` and `right` are two
 * extractors on the same device that have just processed the left and the right images of the same batch (same image size,
 * same parameters); their key points, descriptors and pyramids are consumed where they lie in HBM — no pyramid download.
 * u_right / depth receive mvuRight / mvDepth: entry f*out_stride + i belongs to key point i of left frame f, -1 = no match
 * (slots beyond the frame's key-point count, up to orbgpu_extractor_max_keypoints(), are set to -1 too).  mb / mbf are Frame::mb / Frame::mbf. */
int orbgpu_stereo_matches(orbgpu_extractor* left, orbgpu_extractor* right, float mb, float mbf, float* u_right, float* depth,
                          int out_stride);
/* The same with device output pointers, enqueued on the left extractor's stream (orbgpu_extractor_sync(left) waits). */
int orbgpu_stereo_matches_dev(orbgpu_extractor* left, orbgpu_extractor* right, float mb, float mbf, float* u_right_dev,
                              float* depth_dev, int out_stride);


/* ------------------------------------------------------------------------------------------------
 * Multi-GPU extraction inside one process (the path shards by frame; there is no exchange step and no collective):
 * one host thread + one extractor per device; a call cuts the batch into contiguous frame ranges
 * [g*batch/G, (g+1)*batch/G), every device runs the H2D -> kernels -> D2H pipeline of orbgpu_extract_batch on its range and
 * writes straight into the caller's arrays at the range's offsets (host gather, input order preserved).  The results are
 * byte-identical to orbgpu_extract_batch on one device.  `devices` = n_devices distinct CUDA device indices (NULL = 0..n-1);
 * max_batch_per_device bounds the frames one device works on at a time (longer ranges run as several passes).  Host buffers
 * should be page-locked (cudaHostAlloc / cudaHostRegister) for full copy bandwidth.  The handle is not thread-safe.
 * ---------------------------------------------------------------------------------------------- */
typedef struct orbgpu_multi_extractor orbgpu_multi_extractor;
int orbgpu_multi_extractor_create(orbgpu_multi_extractor** out, const int* devices, int n_devices, int nfeatures, float scale_factor,
                                  int nlevels, int ini_th_fast, int min_th_fast, int max_width, int max_height, int max_batch_per_device);
int orbgpu_multi_extractor_destroy(orbgpu_multi_extractor* me);
int orbgpu_multi_extractor_device_count(const orbgpu_multi_extractor* me);
int orbgpu_multi_extractor_max_keypoints(const orbgpu_multi_extractor* me);
/* Frames [*first, *last) of a batch that device slot g works on. */
int orbgpu_multi_extractor_frame_range(const orbgpu_multi_extractor* me, int batch, int g, int* first, int* last);
/* ORBextractor::operator() for `batch` frames over all devices; arguments as orbgpu_extract_batch (host pointers). */
int orbgpu_multi_extract_batch(orbgpu_multi_extractor* me, const uint8_t* images, int batch, int width, int height, size_t row_stride,
                               size_t frame_stride, orbgpu_keypoint* kp_out, uint8_t* desc_out, int kp_capacity, int32_t* counts);
/* Kernels launched by the last call, summed over the devices. */
int orbgpu_multi_extractor_last_launches(const orbgpu_multi_extractor* me);

/* ------------------------------------------------------------------------------------------------
 * Matching — replaces the Hamming path of ORBmatcher (ORBmatcher.h:37-102)
 *
 * ORBmatcher's search functions read Frame / KeyFrame / MapPoint members; the C ABI takes flat, read-only views of
 * exactly those members, concatenated over a batch of frames, so one call matches many independent frame pairs
 * (the unit the path shards by).  The C++ shell (csrc/host/ORBmatcher.cc) packs one frame per call.
 * ---------------------------------------------------------------------------------------------- */
#define ORBGPU_GRID_COLS 64 /* FRAME_GRID_COLS, Frame.h:38 */
#define ORBGPU_GRID_ROWS 48 /* FRAME_GRID_ROWS, Frame.h:37 */
#define ORBGPU_TH_LOW 50    /* ORBmatcher.cc:38 */
#define ORBGPU_TH_HIGH 100  /* ORBmatcher.cc:37 */
#define ORBGPU_HISTO_LENGTH 30 /* ORBmatcher.cc:39 */

typedef struct orbgpu_matcher orbgpu_matcher;

/* A set of frames (Frame or KeyFrame objects).  Frame f owns keypoints [kp_off[f], kp_off[f+1]). */
typedef struct orbgpu_frame_set {
    int32_t n_frames;
    const int32_t* kp_off;          /* [n_frames+1]                                                              */
    const orbgpu_keypoint* keys_un; /* mvKeysUn: pt, octave, angle are read                                      */
    const uint8_t* desc;            /* mDescriptors rows, 32 contiguous bytes each (ORBmatcher.cc:1903-1904)     */
    const float* u_right;           /* mvuRight; NULL means -1 everywhere (monocular)                            */
    const uint8_t* kp_flags;        /* per keypoint; meaning is stated per search function; NULL means 0         */
    const float* grid;              /* [n_frames][4] mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv
                                       (Frame.cc:101-102; projection search only, else NULL)                      */
    /* DBoW2::FeatureVector of every frame (std::map<NodeId, vector<unsigned>>, FeatureVector.h:21-22), BoW searches
       only: frame f owns nodes [fv_node_off[f], fv_node_off[f+1]), ascending ids; node k owns the feature indices
       fv_feat[fv_feat_off[k] .. fv_feat_off[k+1]) (frame-local keypoint indices, in vector order).  As in DBoW2
       (FeatureVector::addFeature is called once per feature, TemplatedVocabulary.h:1195) a keypoint index appears
       at most once among a frame's nodes. */
    const int32_t* fv_node_off;     /* [n_frames+1]     */
    const int32_t* fv_node_id;      /* [total nodes]    */
    const int32_t* fv_feat_off;     /* [total nodes+1]  */
    const int32_t* fv_feat;         /* [total features] */
} orbgpu_frame_set;

/* The local map points handed to SearchByProjection(Frame&, vector<MapPoint*>&, th) (ORBmatcher.cc:59), per frame:
 * frame f is matched against map points [mp_off[f], mp_off[f+1]) in vector order. */
typedef struct orbgpu_mappoint_set {
    const int32_t* mp_off;     /* [n_frames+1]                                                */
    const float* proj_x;       /* mTrackProjX                                                 */
    const float* proj_y;       /* mTrackProjY                                                 */
    const float* proj_xr;      /* mTrackProjXR (read only where the frame has mvuRight > 0)   */
    const float* view_cos;     /* mTrackViewCos                                               */
    const int32_t* level;      /* mnTrackScaleLevel                                           */
    const uint8_t* flags;      /* bit0 mbTrackInView, bit1 isBad(), bit2 Observations() > 0   */
    const uint8_t* desc;       /* GetDescriptor(), 32 B each                                  */
} orbgpu_mappoint_set;

/* Queries of the generic windowed search: one per projected map point, already reduced to what the search loop reads.
 * Frame f owns queries [q_off[f], q_off[f+1]) in the reference's loop order. */
typedef struct orbgpu_window_query_set {
    const int32_t* q_off;      /* [n_frames+1]                                                                              */
    const float* u;            /* projection, the x / y arguments of Frame::GetFeaturesInArea (Frame.cc:353)                */
    const float* v;
    const float* radius;       /* its r argument (th * mvScaleFactors[level], ORBmatcher.cc:1598, :1760)                    */
    const int32_t* min_level;  /* its minLevel / maxLevel arguments (-1 = unbounded above)                                  */
    const int32_t* max_level;
    const float* ur;           /* u - mbf * invzc, checked against mvuRight of the candidate (:1624-1630); NULL = no check   */
    const uint8_t* flags;      /* bit0: the query is live (passed the projection tests), bit2: its MapPoint has observations */
    const uint8_t* desc;       /* pMP->GetDescriptor(), 32 B each                                                           */
    const float* angle;        /* angle of the query's key point in its own frame (rotation histogram); NULL if unused       */
} orbgpu_window_query_set;

int orbgpu_matcher_create(orbgpu_matcher** out, int device);
int orbgpu_matcher_destroy(orbgpu_matcher* m);
int orbgpu_matcher_sync(orbgpu_matcher* m);
int orbgpu_matcher_stream(orbgpu_matcher* m, void** stream_out);
int orbgpu_matcher_last_launches(const orbgpu_matcher* m);
/* Device time of the last search call (kernels only, CUDA events on the matcher's stream; synchronises) and the
 * number of 256-bit distance evaluations its kernels performed. */
int orbgpu_matcher_last_stats(orbgpu_matcher* m, float* kernel_ms, int64_t* distance_evals);
/* Kernel selection of the BoW-node scans: a node pair with at least min_queries x min_candidates descriptors is
 * scanned by the register-tiled kernel (one thread owns several queries, candidates are staged in shared memory),
 * smaller ones by the warp-per-query kernel.  Defaults 512 / 256; min_queries <= 0 disables the tiled kernel.
 * queries_per_thread (4 or 8, 0 = keep) is the register tile of the tiled kernel.  Results do not depend on any
 * of these. */
int orbgpu_matcher_configure(orbgpu_matcher* m, int min_queries, int min_candidates, int queries_per_thread);

/* Device-resident copies of the views above (key frames live for many searches: upload once, match often).
 * The small per-frame offset arrays are also kept on the host side of the handle. */
typedef struct orbgpu_frame_set_dev orbgpu_frame_set_dev;
typedef struct orbgpu_mappoint_set_dev orbgpu_mappoint_set_dev;
int orbgpu_frame_set_upload(orbgpu_matcher* m, const orbgpu_frame_set* host, orbgpu_frame_set_dev** out);
int orbgpu_frame_set_release(orbgpu_frame_set_dev* fs);
int orbgpu_mappoint_set_upload(orbgpu_matcher* m, const orbgpu_mappoint_set* host, int n_frames, orbgpu_mappoint_set_dev** out);
int orbgpu_mappoint_set_release(orbgpu_mappoint_set_dev* mps);

/* ORBmatcher::DescriptorDistance (ORBmatcher.cc:1901-1917) for n independent descriptor pairs, on the device. */
int orbgpu_hamming_pairs(orbgpu_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* dist_out);

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (ORBmatcher.cc:59-155) for every frame of
 * `frames` against its slice of `mps`.  frames->kp_flags: 0 = mvpMapPoints[idx] is NULL, 1 = it holds a MapPoint
 * with Observations() > 0 (candidate skipped, :108-110), 2 = it holds one with no observations.
 * Outputs (any may be NULL):
 *   kp_match[total keypoints]  index (within the frame's map-point slice) of the map point written into
 *                              F.mvpMapPoints[idx] by this call, -1 where the call wrote nothing
 *   mp_best_idx / mp_best_dist / mp_second_dist [total map points]  bestIdx, bestDist, bestDist2 of every map point
 *                              that reached the selection (:127-139); -1 / 256 / 256 otherwise
 *   nmatches[n_frames]         the function's return value per frame */
int orbgpu_search_by_projection(orbgpu_matcher* m, const orbgpu_frame_set* frames, const orbgpu_mappoint_set* mps,
                                const float* scale_factors, int n_levels, float th, float nnratio, int32_t* kp_match,
                                int32_t* mp_best_idx, int32_t* mp_best_dist, int32_t* mp_second_dist, int32_t* nmatches);

/* The search loop shared by ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (ORBmatcher.cc:1540-1685) and
 * SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist) (:1711-1833), for every frame of `frames`
 * (the CurrentFrame side) against its slice of `queries` (the projected map points; the pose arithmetic that produces them
 * stays in the C++ shell).  Per query: candidates = Frame::GetFeaturesInArea(u, v, radius, min_level, max_level); skip key
 * points holding a MapPoint with observations (kp_flags == 1; with skip_any_mappoint != 0 any MapPoint, kp_flags != 0) and the
 * stereo-inconsistent ones; best = smallest distance, first encountered; accept when best <= th_dist; the key point then
 * holds the query's MapPoint (later queries see that); optional rotation histogram with removal outside the three main bins.
 *   kp_match[total key points]  -1 untouched, >= 0 index (within the frame's query slice) of the MapPoint assigned last,
 *                               -2 reset to NULL by the rotation check (:1676, :1825)
 *   q_best_idx / q_best_dist [total queries] (may be NULL), nmatches[n_frames] the functions' return values */
int orbgpu_search_windowed(orbgpu_matcher* m, const orbgpu_frame_set* frames, const orbgpu_window_query_set* queries, int th_dist,
                           int skip_any_mappoint, int check_orientation, int32_t* kp_match, int32_t* q_best_idx, int32_t* q_best_dist,
                           int32_t* nmatches);

/* ORBmatcher::SearchForTriangulation (ORBmatcher.cc:783-975) for n_pairs keyframe pairs; pair p matches frame
 * idx1[p] of set1 against frame idx2[p] of set2.  kp_flags bit0 = the keypoint already has a MapPoint (skipped,
 * :846,:868).  f12[p] = the 3x3 fundamental matrix, row major; epipole[p] = (ex, ey) as computed at :790-799.
 * match12[match_off[p] + i] receives the index matched to keypoint i of frame idx1[p] or -1 (vMatches12, :964-972);
 * match_dist (may be NULL) the Hamming distance of that match. */
int orbgpu_search_for_triangulation(orbgpu_matcher* m, const orbgpu_frame_set* set1, const orbgpu_frame_set* set2,
                                    int n_pairs, const int32_t* idx1, const int32_t* idx2, const float* f12,
                                    const float* epipole, const float* scale_factors, const float* level_sigma2,
                                    int n_levels, int only_stereo, int check_orientation, const int64_t* match_off,
                                    int32_t* match12, int32_t* match_dist, int32_t* nmatches);

/* The BoW-node scans ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...) (ORBmatcher.cc:635-768; th_inclusive = 0,
 * `bestDist1 < TH_LOW` at :711) and SearchByBoW(KeyFrame*, Frame&, ...) (:211-344; th_inclusive = 1, `<=` at :284).
 * A brute-force scan is the same call with one node holding every index.  kp_flags bit0 = the keypoint has a valid
 * (non-NULL, !isBad) MapPoint: required of the queries of set1, and of the candidates of set2 when
 * require_mp2 != 0 (the KeyFrame-KeyFrame variant, :681-688).  match12 as above (the index in frame 2 whose MapPoint
 * the reference stores in vpMatches12[idx1]); for the KeyFrame-Frame variant the reference indexes its output by the
 * frame-2 keypoint instead — the shell transposes. */
int orbgpu_search_by_bow(orbgpu_matcher* m, const orbgpu_frame_set* set1, const orbgpu_frame_set* set2, int n_pairs,
                         const int32_t* idx1, const int32_t* idx2, float nnratio, int check_orientation, int th_low,
                         int th_inclusive, int require_mp2, const int64_t* match_off, int32_t* match12,
                         int32_t* match_dist, int32_t* nmatches);

/* Device-resident variants: the frame / map-point sets are uploaded handles, every OUTPUT pointer is a device
 * pointer on the matcher's device (NULL allowed where the host variant allows it), the per-pair control arrays
 * (idx1, idx2, match_off, f12, epipole, scale tables) stay host pointers.  The work is enqueued on the matcher's
 * stream and the call returns without synchronising (orbgpu_matcher_sync waits). */
int orbgpu_search_by_projection_dev(orbgpu_matcher* m, const orbgpu_frame_set_dev* frames, const orbgpu_mappoint_set_dev* mps,
                                    const float* scale_factors, int n_levels, float th, float nnratio, int32_t* kp_match_dev,
                                    int32_t* mp_best_idx_dev, int32_t* mp_best_dist_dev, int32_t* mp_second_dist_dev,
                                    int32_t* nmatches_dev);
int orbgpu_search_for_triangulation_dev(orbgpu_matcher* m, const orbgpu_frame_set_dev* set1, const orbgpu_frame_set_dev* set2,
                                        int n_pairs, const int32_t* idx1, const int32_t* idx2, const float* f12,
                                        const float* epipole, const float* scale_factors, const float* level_sigma2,
                                        int n_levels, int only_stereo, int check_orientation, const int64_t* match_off,
                                        int32_t* match12_dev, int32_t* match_dist_dev, int32_t* nmatches_dev);
int orbgpu_search_by_bow_dev(orbgpu_matcher* m, const orbgpu_frame_set_dev* set1, const orbgpu_frame_set_dev* set2, int n_pairs,
                             const int32_t* idx1, const int32_t* idx2, float nnratio, int check_orientation, int th_low,
                             int th_inclusive, int require_mp2, const int64_t* match_off, int32_t* match12_dev,
                             int32_t* match_dist_dev, int32_t* nmatches_dev);

/* Frame::isInFrustum (Frame.cc:274-342) with MapPoint::PredictScale (MapPoint.cc:421-436) for all map points of a batch
 * of frames — what Tracking::SearchLocalPoints runs per local map point before SearchByProjection; the outputs are the
 * mTrack* fields, i.e. exactly the proj_x / proj_y / proj_xr / view_cos / level arrays and bit 0 of `flags` of an
 * orbgpu_mappoint_set.  cam[f] holds 24 floats of frame f: mRcw (9, row-major), mtcw (3), mOw (3), fx, fy, cx, cy, mbf,
 * mnMinX, mnMaxX, mnMinY, mnMaxY.  Frame f owns map points [mp_off[f], mp_off[f+1]): world_pos / normal are xyz triples
 * (GetWorldPos / GetNormal), min_dist_inv / max_dist_inv are GetMinDistanceInvariance() / GetMaxDistanceInvariance(),
 * max_distance is mfMaxDistance (PredictScale's numerator).  Points failing a test get in_view 0 and zeroed outputs.
 * The cv::Mat arithmetic is evaluated as OpenCV 4.13 does (float dot product + double addend in cv::gemm, double norm and
 * dot); the predicted level uses a correctly rounded logf — it can differ from glibc's logf by one level only when
 * log(ratio)/log(scaleFactor) is within one ulp of an integer. */
int orbgpu_is_in_frustum(orbgpu_matcher* m, int n_frames, const float* cam, float log_scale_factor, int n_levels,
                         float viewing_cos_limit, const int32_t* mp_off, const float* world_pos, const float* normal,
                         const float* min_dist_inv, const float* max_dist_inv, const float* max_distance, uint8_t* in_view,
                         float* proj_x, float* proj_y, float* proj_xr, int32_t* level, float* view_cos);
/* The same with every array but `mp_off` and `cam` (host) a device pointer; enqueued on the matcher's stream. */
int orbgpu_is_in_frustum_dev(orbgpu_matcher* m, int n_frames, const float* cam, float log_scale_factor, int n_levels,
                             float viewing_cos_limit, const int32_t* mp_off, const float* world_pos_dev, const float* normal_dev,
                             const float* min_dist_inv_dev, const float* max_dist_inv_dev, const float* max_distance_dev,
                             uint8_t* in_view_dev, float* proj_x_dev, float* proj_y_dev, float* proj_xr_dev, int32_t* level_dev,
                             float* view_cos_dev);

/* The projection step of Tracking::SearchLocalPoints (Tracking.cc:1150-1200) straight into a device-resident map-point set:
 * uploads the local map (same inputs as orbgpu_is_in_frustum, plus `flags` — bit 1: isBad(), bit 2: Observations() > 0, bit 0
 * is ignored — and the descriptors), runs isInFrustum on the device and leaves proj_x / proj_y / proj_xr / view_cos / level and
 * the visibility bit where orbgpu_search_by_projection_dev reads them: projection -> window -> candidates -> match without
 * the mTrack* fields ever visiting the host.  Release with orbgpu_mappoint_set_release. */
int orbgpu_mappoint_set_project(orbgpu_matcher* m, int n_frames, const float* cam, float log_scale_factor, int n_levels,
                                float viewing_cos_limit, const int32_t* mp_off, const float* world_pos, const float* normal,
                                const float* min_dist_inv, const float* max_dist_inv, const float* max_distance, const uint8_t* flags,
                                const uint8_t* desc, orbgpu_mappoint_set_dev** out);

/* Best-only windowed search: the candidate loops of ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th)
 * (ORBmatcher.cc:1051-1112), Fuse(KeyFrame*, Scw, ...) (:1211-1246) and both directions of SearchBySim3 (:1363-1401,
 * :1443-1481).  Queries are independent — no key point is taken by an earlier query; the map updates that follow in the
 * reference (Replace / AddObservation / the agreement check) stay with the caller.  For every live query (flags bit 0):
 * candidates = key points of the frame inside the window (u, v, radius) with octave in [min_level, max_level]
 * (KeyFrame::GetFeaturesInArea order, KeyFrame.cc:583-622), minus those with kp_flags != 0 when skip_flagged
 * (vbAlreadyMatched2 / vbAlreadyMatched1 of SearchBySim3); with inv_level_sigma2 != NULL (mvInvLevelSigma2, n_levels entries)
 * a candidate must also pass Fuse's chi-square gate: e2 * invSigma2[octave] <= 5.99 with e2 = ex^2 + ey^2 for a monocular
 * key point (mvuRight < 0), <= 7.8 with e2 = ex^2 + ey^2 + er^2 for a stereo one (er from queries->ur, which must be given).
 * Outputs: the candidate of least distance (first wins) and that distance; -1 / 256 if none.  The caller compares with
 * TH_LOW (Fuse) or TH_HIGH (SearchBySim3). */
int orbgpu_search_window_best(orbgpu_matcher* m, const orbgpu_frame_set* frames, const orbgpu_window_query_set* queries,
                              const float* inv_level_sigma2, int n_levels, int skip_flagged, int32_t* q_best_idx,
                              int32_t* q_best_dist);

/* ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vbPrevMatched, vnMatches12, windowSize) (ORBmatcher.cc:493-632),
 * the monocular bootstrap search, for a batch of frame pairs.  frames2 = the F2 of every pair (with grid); queries1 = the key
 * points of the matching F1 as window queries: u / v = vbPrevMatched[i1], radius = windowSize, min_level = max_level = 0,
 * flags bit 0 = (octave == 0) (:512-514), desc / angle = F1's descriptor row / key-point angle.  match12[q] = vnMatches12
 * (index in F2 or -1) after the distance test (<= TH_LOW), the ratio test (bestDist < bestDist2 * nnratio), the stealing rule
 * (a better match takes a key point from its earlier owner, :546, :573-583) and the rotation histogram; nmatches per pair.
 * The caller updates vbPrevMatched from match12 (:626-628). */
int orbgpu_search_for_initialization(orbgpu_matcher* m, const orbgpu_frame_set* frames2, const orbgpu_window_query_set* queries1,
                                     float nnratio, int check_orientation, int32_t* match12, int32_t* nmatches);

/* MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:247-316) for a batch of map points: point p owns the descriptors
 * [obs_off[p], obs_off[p+1]) of `desc` (the rows of its observing, non-bad key frames in the reference's iteration order,
 * 32 bytes each, at most 256 per point).  best_idx[p] = the row (relative to obs_off[p]) with the least median Hamming
 * distance to all rows, its own included — the median is the element of rank (int)(0.5*(N-1)) of the sorted row, the first
 * row wins ties; -1 for a point without descriptors.  best_median (may be NULL) receives that median. */
int orbgpu_distinctive_descriptors(orbgpu_matcher* m, int n_points, const int32_t* obs_off, const uint8_t* desc, int32_t* best_idx,
                                   int32_t* best_median);

/* ------------------------------------------------------------------------------------------------
 * Vocabulary — replaces ORBVocabulary::transform (DBoW2 TemplatedVocabulary<FORB::TDescriptor, FORB>,
 * include/ORBVocabulary.h:31-32) as called by Frame::ComputeBoW (Frame.cc:425-432) and KeyFrame::ComputeBoW
 * (KeyFrame.cc:59-70): every descriptor descends the vocabulary tree by FORB::distance (FORB.cpp:81-101, first minimum
 * wins, TemplatedVocabulary.h:1237-1247); the words reached form the BowVector, the nodes passed `levelsup` levels above
 * the leaves group the feature indices into the FeatureVector the BoW searches consume.
 * ---------------------------------------------------------------------------------------------- */
typedef struct orbgpu_vocabulary orbgpu_vocabulary;

/* The vocabulary as TemplatedVocabulary::loadFromTextFile reads it (TemplatedVocabulary.h:1338-1423): k, L, scoring
 * (ScoringType, BowVector.h:45-53: 0 L1_NORM .. 5 DOT_PRODUCT) and weighting (WeightingType, BowVector.h:36-42: 0 TF_IDF,
 * 1 TF, 2 IDF, 3 BINARY) from the header line, then one record per line: record r describes node r + 1 (node 0 is the
 * root) — parent[r] (< r + 1), is_leaf[r], desc[r][32], weight[r].  Word ids are handed out to the leaves in record
 * order (:1407-1414).  At most 32 children per node. */
int orbgpu_vocabulary_create(orbgpu_vocabulary** out, int device, int k, int L, int scoring, int weighting, int n_records,
                             const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weight);
int orbgpu_vocabulary_destroy(orbgpu_vocabulary* v);
/* Thread safety.  The reference shares ONE ORBVocabulary between its threads (Tracking calls Frame::ComputeBoW, Tracking.cc:874,
 * :1585, while LocalMapping calls KeyFrame::ComputeBoW, LocalMapping.cc:164).  A handle owns one stream and one set of scratch
 * buffers: orbgpu_bow_transform (host pointers) locks the handle for the whole call, so concurrent callers of one handle are
 * serialised, never interleaved; orbgpu_bow_transform_dev does not lock (one thread per handle).  For real concurrency give
 * every host thread its own fork: a second handle on the SAME read-only tree in HBM (no copy; the tree is freed with the last
 * handle on it) with its own stream and scratch.  Destroy a fork with orbgpu_vocabulary_destroy, in any order. */
int orbgpu_vocabulary_fork(const orbgpu_vocabulary* v, orbgpu_vocabulary** out);
int orbgpu_vocabulary_info(const orbgpu_vocabulary* v, int* n_nodes, int* n_words);
int orbgpu_vocabulary_sync(orbgpu_vocabulary* v);
int orbgpu_vocabulary_last_launches(const orbgpu_vocabulary* v);
/* The CUDA stream (cudaStream_t) the vocabulary's work is enqueued on, for callers that order their own work after it. */
int orbgpu_vocabulary_stream(orbgpu_vocabulary* v, void** stream_out);

/* transform(features, BowVector&, FeatureVector&, levelsup) (TemplatedVocabulary.h:1127-1197) for every frame of a batch:
 * frame f owns descriptor rows [kp_off[f], kp_off[f+1]) of `desc`.  No size limit, like the reference: one CTA sorts a frame's
 * (node, index) and (word, index) keys in shared memory while the largest frame of the batch has at most 8192 descriptors
 * (the reference's settings stay far below: nFeatures 1000-2000, 2 x nFeatures for the monocular initialiser); a batch with a
 * larger frame runs the same kernel on a block-private workspace in HBM (slower, same results).  All outputs are sized by the
 * caller for n = kp_off[n_frames] entries (the +1 arrays for n + 1) and may be NULL when not wanted:
 *   BowVector     (std::map<WordId, WordValue>, BowVector.h:58-60): frame f owns entries [bv_off[f], bv_off[f+1]) of
 *                 bv_word (ascending) / bv_value (after addWeight / addIfNotExist and normalisation, BowVector.cpp:36-90,
 *                 bit-exact doubles: the sums run in the reference's order);
 *   FeatureVector (std::map<NodeId, vector<unsigned>>, FeatureVector.h:21-22) in the orbgpu_frame_set layout: frame f owns
 *                 nodes [fv_node_off[f], fv_node_off[f+1]); node j has id fv_node_id[j] (ascending within the frame) and the
 *                 frame-local feature indices fv_feat[fv_feat_off[j] .. fv_feat_off[j+1]) in ascending order;
 *   word_of_feature / node_of_feature: the WordId and the level-(L - levelsup) NodeId of every descriptor
 *                 (transform(feature, id, weight, &nid, levelsup), :1214-1259).
 * Features whose word has weight 0 (stopped words, :1185) appear in neither vector.  A leaf shallower than level
 * L - levelsup (where the reference leaves nid uninitialised) reports its own node id. */
int orbgpu_bow_transform(orbgpu_vocabulary* v, int n_frames, const int32_t* kp_off, const uint8_t* desc, int levelsup,
                         int32_t* bv_off, uint32_t* bv_word, double* bv_value, int32_t* fv_node_off, int32_t* fv_node_id,
                         int32_t* fv_feat_off, int32_t* fv_feat, uint32_t* word_of_feature, uint32_t* node_of_feature);
/* The same with `kp_off_dev`, `desc_dev` and every output a device pointer on the vocabulary's device (n_features = the host's
 * copy of kp_off[n_frames], max_per_frame = an upper bound of the rows per frame); enqueued on the vocabulary's stream. */
int orbgpu_bow_transform_dev(orbgpu_vocabulary* v, int n_frames, const int32_t* kp_off_dev, int n_features, int max_per_frame,
                             const uint8_t* desc_dev, int levelsup, int32_t* bv_off, uint32_t* bv_word, double* bv_value,
                             int32_t* fv_node_off, int32_t* fv_node_id, int32_t* fv_feat_off, int32_t* fv_feat,
                             uint32_t* word_of_feature, uint32_t* node_of_feature);

/* ------------------------------------------------------------------------------------------------
 * Chaining on the device: extraction -> vocabulary -> matching without a host round trip.
 * Builds a device-resident frame set from the batch `ex` has just processed (its key points and descriptors are compacted
 * where they lie in HBM; they serve as mvKeysUn, i.e. an undistorted / rectified camera as in the KITTI and EuRoC-rectified
 * configurations — with distortion, undistort on the host and use orbgpu_frame_set_upload).  With `voc` != NULL every frame's
 * FeatureVector is computed by orbgpu_bow_transform_dev (levelsup as in Frame::ComputeBoW) and attached.  kp_flag fills
 * kp_flags (meaning per search function); u_right_dev (optional, device, [batch][u_right_stride], e.g. the output of
 * orbgpu_stereo_matches_dev) becomes mvuRight; grid (optional, host, 4 floats: mnMinX, mnMinY, mfGridElementWidthInv,
 * mfGridElementHeightInv) is shared by all frames.  Only the per-frame counts travel to the host.  Release with
 * orbgpu_frame_set_release BEFORE destroying `m`: the set's memory is borrowed from the matcher and goes back to it, so a
 * pipeline that builds one set per batch allocates nothing after the first step.  Use such a set only with `m`. */
int orbgpu_frame_set_from_extraction(orbgpu_matcher* m, orbgpu_extractor* ex, orbgpu_vocabulary* voc, int levelsup, int kp_flag,
                                     const float* u_right_dev, int u_right_stride, const float* grid, orbgpu_frame_set_dev** out);
/* Sizes of a device-resident frame set; kp_off (may be NULL) receives n_frames + 1 offsets. */
int orbgpu_frame_set_dev_info(const orbgpu_frame_set_dev* fs, int* n_frames, int32_t* kp_off, int* n_nodes, int* n_feat);
/* Copies arrays of a device-resident frame set to the host (any pointer may be NULL; sizes from orbgpu_frame_set_dev_info). */
int orbgpu_frame_set_download(orbgpu_matcher* m, const orbgpu_frame_set_dev* fs, orbgpu_keypoint* keys, uint8_t* desc,
                              int32_t* fv_node_off, int32_t* fv_node_id, int32_t* fv_feat_off, int32_t* fv_feat);

#ifdef __cplusplus
}
#endif
#endif /* ORBGPU_H */
