// og_types.h — device-visible geometry of one extractor instance (plain structs, passed to kernels by value
// or through device tables).  All of it is derived on the host exactly as the reference constructor and
// ComputePyramid / ComputeKeyPointsOctTree derive it (ORBextractor.cc:410-470, :1107-1116, :765-806).
#pragma once
#include <stdint.h>

#include "og_math.cuh"

namespace og {

constexpr int kMaxLevels = 12;
constexpr int kEdge = 19;        // EDGE_THRESHOLD (ORBextractor.cc:74): border kept around every level
constexpr int kXPad = 32;        // the level interior starts at byte column kXPad (>= kEdge, keeps rows 32-B aligned)
constexpr int kHalfPatch = 15;   // HALF_PATCH_SIZE (:73)
constexpr int kCellMax = 64;     // upper bound of a cell's tested width/height (wCell < 60 by construction)

// HBM layout of the pyramid: pyr[level][frame][rows][pitch] bytes, interior pixel (x,y) of a level at
// row kEdge + y, column kXPad + x.  The 19-px BORDER_REFLECT_101 frame of the reference's level buffers
// (:1113-1128) occupies columns [kXPad-19, kXPad+w+19) and rows [0, h+38).
// One output index of cv::resize's coefficient tables: two source indices and two 11-bit weights.
struct Tap {
    int16_t s0, s1, w0, w1;
};

struct Level {
    const Tap* xt;           // [w] horizontal taps from level l-1 (null for level 0)
    const Tap* yt;           // [h] vertical taps
    const uint4* ytw;        // [h] the same for k_resize_tma, decoded: {s0 * 256, s1 * 256, w0 << 16, w1 << 16} (one 16-byte load, no unpacking)
    int w, h;                // cvRound(cols * mvInvScaleFactor[l]) (:1112)
    int pitch, rows;         // bytes per row (multiple of 128), h + 2*kEdge
    long long frame_stride;  // pitch * rows
    long long base;          // byte offset of [level][frame 0]
    // detection grid (:770-806)
    int cell_base, n_cells;  // this level's slice of the cell table
    int cand_base, cand_cap; // this level's slice of a frame's candidate slots (sum of per-cell bounds)
    int quota;               // mnFeaturesPerLevel[l] (:435-446)
    int n_ini;               // round(width/height) root nodes (:543)
    float hx;                // root width (:545)
    int det_w, det_h;        // maxBorderX-minBorderX, maxBorderY-minBorderY
    int node_cap;            // node array capacity of the octree (max(quota+3, 4*n_ini) + slack)
    int sel_base, sel_cap;   // this level's slice of a frame's selected-keypoint slots
    float scale;             // mvScaleFactor[l]
    float kp_size;           // (float)(int)(31 * scale) (:837)
    long long ot_base;       // byte offset of this level's octree workspace inside a frame's workspace
    int wcell, hbox;         // cell pitch in x (:785); rows of the FAST tile box (tallest cell + 6)
    int wcell_magic, pad_;   // ceil(2^16 / wcell): px / wcell = (px * wcell_magic) >> 16 for px < 256
};

// One FAST cell (:789-806): the reference runs cv::FAST on the sub-image [ini, max) and FAST itself never tests
// the outer 3 px, so the tested ("owned") pixels are [x0, x0+tw) x [y0, y0+th) in level coordinates.
struct Cell {
    int16_t level;
    int16_t x0, y0;   // iniX+3, iniY+3
    int16_t tw, th;   // tested width / height (> 0)
    int16_t pad;
    int32_t slot;     // first candidate slot of this cell inside the level's slice
};

// A run of consecutive cells of one cell row whose tested pixels fit one 256-byte-wide shared-memory tile
// (k_fast_seg): first tested pixel (x0, y0) in level coordinates, tested extent tw x th.
constexpr int kSegPitch = 256;   // tile row pitch in bytes = TMA box width
constexpr int kSegMaxTw = 232;   // TMA starts at a 16-byte aligned column: up to 18 bytes before the first tested pixel, 3 after the last
struct Segment {
    int16_t level, ncells;
    int32_t first_cell;          // index into the frame's cell table
    int16_t x0, y0, tw, th;
    uint32_t nw_magic;           // ceil(2^32 / nb), nb = 8-pixel steps covering the tested columns: row = (unit * nw_magic) >> 32
};

struct ExtractParams {
    Level lv[kMaxLevels];
    int n_levels;
    int batch;                // frames of this launch
    int frame0;               // first frame of this launch inside the workspace / output arrays (chunked host path)
    int ini_th, min_th;
    int total_cells;          // cells per frame (all levels)
    int total_cand_cap;       // candidate slots per frame
    int total_sel_cap;        // selected-keypoint slots per frame
    long long ot_frame_bytes; // octree workspace per frame
    int kp_cap;               // output capacity per frame
    AtanCoef atan;
    float factor_pi;          // (float)(CV_PI/180.f) (:107)
    int umax[16];             // (:454-469)
    const uint32_t* ic_tab;   // IC_Angle DP4A weights [4 alignments][31 rows][9 words][2] (k_orient_desc)
    // device buffers
    uint8_t* pyr;
    uint8_t* blur;
    const Cell* cells;
    const Segment* segs;
    int n_segs;
    int32_t* cell_count;      // [batch][total_cells]
    uint32_t* cand_xy;        // [batch][total_cand_cap]   y<<16 | x, relative to (minBorderX, minBorderY)
    uint8_t* cand_resp;       // [batch][total_cand_cap]
    uint8_t* ot_ws;           // [batch][ot_frame_bytes]
    uint32_t* sel_xy;         // [batch][total_sel_cap]
    uint8_t* sel_resp;        // [batch][total_sel_cap]
    int32_t* sel_count;       // [batch][n_levels]
};

struct KeyPoint {  // == cv::KeyPoint == orbgpu_keypoint
    float x, y, size, angle, response;
    int32_t octave, class_id;
};

}  // namespace og
