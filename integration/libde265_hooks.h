// libde265_hooks.h — the reference-side binding for the B2 boundary.
//
// These are the only declarations a libde265 maintainer adds to the tree; the five call sites are
//   slice.cc   decode_TU()                          -> b200_hook_decode_TU
//   slice.cc   read_pcm_samples_internal()          -> b200_hook_pcm
//   motion.cc  generate_inter_prediction_samples()  -> b200_hook_inter_pred
//   decctx.cc  decode_some() post-processing        -> b200_hook_picture_done
//   decctx.h   base_context                         -> void* b200_state
// oracle/patch_ref.py applies exactly these insertions to a scratch copy of the reference for the tests.
#ifndef LIBDE265_B200_HOOKS_H
#define LIBDE265_B200_HOOKS_H

#include <stddef.h>
#include <stdint.h>

#include "b200hevc.h"

class thread_context;
class base_context;
class decoder_context;
class slice_segment_header;
struct de265_image;
class PBMotion;

// Further call sites of the built-in backend (de265_acceleration_B200):
//   de265.h    enum de265_acceleration                    -> de265_acceleration_B200 = 200
//   decctx.cc  base_context::set_acceleration_functions   -> b200_hook_set_acceleration   (DE265_DECODER_PARAM_ACCELERATION_CODE, de265.cc:576-578)
//   decctx.cc  generate_unavailable_reference_picture     -> b200_hook_unavailable_reference
//   de265.cc   de265_peek_next_picture                    -> b200_hook_wait_image         (the deferred read-back is awaited when the picture is handed out)
// true => the hook consumed the call and the host must NOT reconstruct.
bool b200_hook_decode_TU(thread_context* tctx, int x0, int y0, int nT, int cIdx, int cuPredMode, bool cbf);
bool b200_hook_inter_pred(base_context* ctx, const slice_segment_header* shdr, de265_image* img, int xP, int yP, int nPbW, int nPbH,
                          const PBMotion* vi);
void b200_hook_pcm(thread_context* tctx, int x0, int y0, int w, int h, int cIdx);
bool b200_hook_picture_done(decoder_context* ctx, de265_image* img);
void b200_hook_set_acceleration(base_context* ctx, int level);
void b200_hook_unavailable_reference(decoder_context* ctx, de265_image* img);
void b200_hook_wait_image(decoder_context* ctx, const de265_image* img);

#define DE265_ACCELERATION_B200 200  /* value of de265_acceleration_B200 in the patched de265.h */

extern "C" {
// A sink receives every finished picture's command records and must leave the final (post-SAO) samples in the given host
// planes (strides in bytes) before returning 0 — or return DE265_B200_SINK_PENDING after STARTING the transfer: the decoder
// then calls the wait callback (de265_b200_set_wait) for that DPB slot before it hands the picture to the application.
// A negative return marks the picture as damaged (integrity = INTEGRITY_DECODING_ERRORS).
typedef int (*de265_b200_sink)(void* user, const b200_picture* pic, void* const planes[3], const size_t strides[3]);
#define DE265_B200_SINK_PENDING 1
typedef int (*de265_b200_wait)(void* user, int dpb_slot);
typedef int (*de265_b200_fill)(void* user, int dpb_slot, const b200_pic_params* params, int value_y, int value_c);
// Attach to a decoder created with de265_new_decoder(); sink==NULL detaches.  Returns 0, or a negative B200_ERR_* code when the
// decoder cannot be served (worker threads running: the recorder is single-threaded).
B200_API int de265_b200_attach(void* de265_decoder_ctx, de265_b200_sink sink, void* user);
// Optional callbacks: `wait` for sinks that return DE265_B200_SINK_PENDING; `fill` mirrors libde265's synthesised reference
// pictures (generate_unavailable_reference_picture, decctx.cc:1294) into the backend.
B200_API void de265_b200_set_callbacks(void* de265_decoder_ctx, de265_b200_wait wait, de265_b200_fill fill);
// The built-in backend: a B200 engine on `device`, an asynchronous sink (records submitted at picture end, read-back into
// page-locked picture planes, awaited when the picture is output).  This is what
//   de265_set_parameter_int(ctx, DE265_DECODER_PARAM_ACCELERATION_CODE, de265_acceleration_B200)
// selects.  Returns 0 or a negative B200_ERR_* code (no CUDA device: B200_ERR_NO_DEVICE — there is no CPU fallback).
B200_API int de265_b200_enable(void* de265_decoder_ctx, int device);
B200_API void de265_b200_disable(void* de265_decoder_ctx);
}

#endif
