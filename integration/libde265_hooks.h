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

// true => the hook consumed the call and the host must NOT reconstruct.
bool b200_hook_decode_TU(thread_context* tctx, int x0, int y0, int nT, int cIdx, int cuPredMode, bool cbf);
bool b200_hook_inter_pred(base_context* ctx, const slice_segment_header* shdr, de265_image* img, int xP, int yP, int nPbW, int nPbH,
                          const PBMotion* vi);
void b200_hook_pcm(thread_context* tctx, int x0, int y0, int w, int h, int cIdx);
bool b200_hook_picture_done(decoder_context* ctx, de265_image* img);

extern "C" {
// A sink receives every finished picture's command records and must leave the final
// (post-SAO) samples in the given host planes (strides in bytes) before returning.
typedef int (*de265_b200_sink)(void* user, const b200_picture* pic, void* const planes[3], const size_t strides[3]);
// Attach to a decoder created with de265_new_decoder(); sink==NULL detaches.
B200_API void de265_b200_attach(void* de265_decoder_ctx, de265_b200_sink sink, void* user);
}

#endif
