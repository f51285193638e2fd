#include "nct_internal.h"
extern "C" {
__attribute__((weak)) void nct_pipe_free(nct_ctx *) {}
}
