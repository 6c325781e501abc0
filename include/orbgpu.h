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

#ifdef __cplusplus
}
#endif
#endif /* ORBGPU_H */
