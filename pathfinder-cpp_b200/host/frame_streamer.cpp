#include "frame_streamer.h"

#include <chrono>
#include <vector>

namespace {

int submit_one(pfcu_ctx *c, const pfhost_scene &s, uint8_t *pixels) {
    int r;
    for (int which = 0; which < 2; which++)
        if ((r = pfcu_upload_scene(c, which, s.points[which], s.n_points[which], s.indices[which], s.n_segments[which]))) return r;
    if ((r = pfcu_begin_frame(c))) return r;
    for (uint32_t i = s.n_clip_batches; i-- > 0;)
        if (s.clip_batches[i].path_count > 0 && (r = pfcu_prepare_batch(c, &s.clip_batches[i]))) return r;
    static const float zero[4] = {0.f, 0.f, 0.f, 0.f};
    bool first = true;
    for (uint32_t i = 0; i < s.n_draw_batches; i++) {
        if ((r = pfcu_prepare_batch(c, &s.draw_batches[i]))) return r;
        const pfhost_draw &d = s.draws[i];
        if (d.target_page < 0) {
            r = pfcu_draw_batch(c, d.batch_id, -1, d.color_page, d.sampling_flags, first ? 1 : 0, s.clear_color);
            first = false;
        } else {
            r = pfcu_draw_batch(c, d.batch_id, d.target_page, d.color_page, d.sampling_flags, 1, zero);
        }
        if (r) return r;
    }
    if ((r = pfcu_submit_frame(c))) return r;
    if (pixels && (r = pfcu_read_target_async(c, pixels, 0))) return r;
    return PFCU_OK;
}

}  // namespace

extern "C" int pfhost_stream_frames(pfcu_ctx *const *contexts, uint32_t n_contexts, const pfhost_scene *scene, uint32_t n_frames,
                                    uint8_t *const *pixels, double *wall_seconds, pfcu_frame_stats *last, uint32_t *retries) {
    if (!contexts || !n_contexts || !scene) return PFCU_ERR_INVALID;
    std::vector<char> pending(n_contexts, 0);
    pfcu_frame_stats st = {};
    uint32_t n_retries = 0;
    const auto t0 = std::chrono::steady_clock::now();
    auto collect = [&](uint32_t k) -> int {
        int r = pfcu_wait_frame(contexts[k], &st);
        if (r) return r;
        n_retries += st.retries;
        if (pixels && (r = pfcu_wait_read(contexts[k]))) return r;
        pending[k] = 0;
        return PFCU_OK;
    };
    // on an error: the frames still in flight are waited for (their status is dropped), so that every context is usable again
    auto bail = [&](int err) {
        for (uint32_t k = 0; k < n_contexts; k++)
            if (pending[k]) {
                pfcu_wait_frame(contexts[k], nullptr);
                if (pixels) pfcu_wait_read(contexts[k]);
                pending[k] = 0;
            }
        return err;
    };
    for (uint32_t i = 0; i < n_frames; i++) {
        const uint32_t k = i % n_contexts;
        int r;
        if (pending[k] && (r = collect(k))) return bail(r);
        if ((r = submit_one(contexts[k], *scene, pixels ? pixels[k] : nullptr))) return bail(r);
        pending[k] = 1;
    }
    // collect in submission order, so that `last` is the last frame's
    for (uint32_t j = 0; j < n_contexts; j++) {
        const uint32_t k = (n_frames + j) % n_contexts;
        int r;
        if (pending[k] && (r = collect(k))) return bail(r);
    }
    if (wall_seconds) *wall_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (last) *last = st;
    if (retries) *retries = n_retries;
    return PFCU_OK;
}
