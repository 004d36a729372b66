// Shim with the reference's header name (/root/reference/Auncel/profile.h) so that the reference's own
// drivers (eval/bound.cpp, eval/effect_error.cpp, dist/worker.cpp) compile unmodified against the
// B200 library: everything they use from this header is declared in auncel/faiss_api.h.
#pragma once
#include <auncel/faiss_api.h>  // compile with -I <repo>/include
