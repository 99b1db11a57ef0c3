/* Plain-C restatement of the codebook delay / revert gathers.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Built by oracle/Makefile
 * into oracle/_build/libdelay_ref.so and loaded by tests through ctypes.
 *
 * Follows the reference's dia/audio.py:
 *   apply   : build_delay_indices  :6-41  + apply_audio_delay  :44-85
 *   revert  : build_revert_indices :88-122 + revert_audio_delay :125-163
 * Pinned against the numpy oracle, the reference itself (validate script) and
 * tests/golden/delay_known_answer.json.
 */
#include <stdint.h>

/* out[b,t,c] = bos            if t - delay[c] <  0
 *            = pad            if t - delay[c] >= T   (cannot happen, kept for fidelity)
 *            = in[b, t-delay[c], c] otherwise                                   */
int delay_ref_apply_i32(const int32_t *in, int32_t *out, int B, int T, int C,
                        const int32_t *delay, int32_t pad, int32_t bos)
{
    if (B < 0 || T < 0 || C < 0) return -1;
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < T; ++t)
            for (int c = 0; c < C; ++c) {
                int ti = t - delay[c];
                int tc = ti < 0 ? 0 : (ti > T - 1 ? T - 1 : ti);
                int32_t g = in[((int64_t)b * T + tc) * C + c];
                out[((int64_t)b * T + t) * C + c] = ti < 0 ? bos : (ti >= T ? pad : g);
            }
    return 0;
}

/* out[b,t,c] = in[b, min(t+delay[c], T-1), c]; PAD where the clamped index >= T_orig */
int delay_ref_revert_i32(const int32_t *in, int32_t *out, int B, int T, int C,
                         const int32_t *delay, int32_t pad, int T_orig)
{
    if (B < 0 || T < 0 || C < 0) return -1;
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < T; ++t)
            for (int c = 0; c < C; ++c) {
                int ti = t + delay[c];
                if (ti > T - 1) ti = T - 1;
                int32_t g = in[((int64_t)b * T + ti) * C + c];
                out[((int64_t)b * T + t) * C + c] = ti >= T_orig ? pad : g;
            }
    return 0;
}
