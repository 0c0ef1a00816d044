// vt_scale_plan.cuh -- the scaling plan shared by the scaler's translation units.
#pragma once
#include <vector>

#include "vt_common.cuh"

struct vt_scale_plan {
    int sw, sh, dw, dh, flags;
    int csw, csh, cdw, cdh;  // chroma plane sizes
    // filter banks in device memory; index 0 = luma, 1 = chroma
    int htaps[2], vtaps[2];
    int16_t *hcoef[2];  // dw x htaps
    int32_t *hpos[2];
    int16_t *vcoef[2];  // dh x vtaps
    int32_t *vpos[2];
    int16_t *scratch;  // generic path: scratch_frames x (dw x sh) int16
    int scratch_frames;
    // fast two-pass kernels (planes the pair kernel does not take): coefficient PAIRS per output sample padded to hp2
    // pairs, vertical bank padded to vt2 taps; 0 = outside the instantiated range (general kernels)
    int hp2[2], vt2[2];
    uint32_t *hc2[2];
    int16_t *vc2[2];
    // host copies (table construction for the pair kernel)
    std::vector<int16_t> h_hcoef[2], h_vcoef[2];
    std::vector<int32_t> h_hpos[2], h_vpos[2];
    // pair kernel (vt_scale_pair.cu), one per plane kind: tables built by vt::build_pair() at plan creation
    struct Pair {
        bool ok = false;
        int hp = 0, tv = 0;            // dp2a pairs of the even column (the odd one uses hp+1), padded vertical taps
        int np = 0;                    // column pairs per lane (2 luma, 1 chroma)
        int strip_cols = 0, n_strips = 0;
        int tile_w = 0, n_boxes = 0, groups_per_stage = 0, stage_rows = 0;
        int box_bytes = 0, stage_bytes = 0, n_stages = 0, warp_smem = 0;
        int lt_words = 0;              // words per lane-table entry
        int32_t *box_x0 = nullptr;     // n_strips x n_boxes: first source byte of each TMA box (multiple of 16)
        int32_t *strip_col = nullptr;  // n_strips: first output column of each strip
        uint32_t *lane_tab = nullptr;  // (n_strips*np*32) x lt_words
        std::vector<int32_t> vtab;     // dh x vstride: front-padded coefficients, then the window's last source row
        int vstride = 0;
        double src_rows_per_dst_row = 1.0;
        // static vertical schedule (0 = none): see find_static_schedule()
        int mask = 0, n_phases = 1, align_p = 1, align_r0 = 0, reg_lo = 0, reg_hi = 0;
        int sc[24] = {0};
        // static horizontal pattern (exact 3:2): a lane owns 8 ADJACENT output columns (luma) / 4 (chroma); every
        // column's tap alignment is then a compile-time constant.  Tables beside the generic ones; see build_pair()
        bool hs = false;
        bool score_ok = false;         // the luma kernel can also produce SAD + histogram of the source (fused K3)
        int32_t *box_x0_hs = nullptr;  // n_strips: first source byte of the strip's single TMA box (multiple of 4, may be < 0)
        uint32_t *lane_tab_hs = nullptr;  // (n_strips*32) x (columns per lane * hp) coefficient pairs
    } pair[2];
};

namespace vt {
int build_pair(vt_scale_plan *p, int c);
void free_pair(vt_scale_plan *p);
struct PairScore {                     // outputs of the fused scene score; sad/hist must be zeroed before the launch
    const uint8_t *prev0;
    uint64_t *sad;
    uint32_t *hist;
};
int launch_pair(const vt_scale_plan *p, int c, const uint8_t *src, int pitch, size_t src_fs, uint8_t *dst, size_t dst_fs,
                int n_frames, cudaStream_t st, const PairScore *score = nullptr);
int launch_score(const uint8_t *luma, int pitch, size_t frame_stride, int w, int h, const uint8_t *prev0, int n_frames,
                 uint64_t *sad, uint32_t *hist, cudaStream_t st);
int make_tmap_u32_3d(void *tmap_out, const uint8_t *base, int row_bytes, int rows, int n_frames, int pitch,
                     size_t frame_stride, int tile_w, int tile_h);
int make_tmap_u8_3d(void *tmap_out, const uint8_t *base, int row_bytes, int rows, int n_frames, int pitch,
                    size_t frame_stride, int tile_w, int tile_h);
}  // namespace vt

