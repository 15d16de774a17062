// eval_tile_kernel (bear_eval.cuh) instantiated for the head variants BEAR_HEAD_NONE, BEAR_HEAD_STOP, BEAR_HEAD_EXPLICIT.
#define BEAR_EVAL_IMPL
#include "bear_eval.cuh"

int bear_eval::launch_misc(const EvalArgs& a) {
    switch (a.head) {
        case BEAR_HEAD_NONE: return launch_head<BEAR_HEAD_NONE>(a);
        case BEAR_HEAD_STOP: return launch_head<BEAR_HEAD_STOP>(a);
        case BEAR_HEAD_EXPLICIT: return launch_head<BEAR_HEAD_EXPLICIT>(a);
    }
    return BEAR_ERR_ARG;
}
