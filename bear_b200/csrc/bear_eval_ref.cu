// eval_tile_kernel (bear_eval.cuh) instantiated for the head variants BEAR_HEAD_REF_STOP, BEAR_HEAD_REF_LINEAR.
#define BEAR_EVAL_IMPL
#include "bear_eval.cuh"

int bear_eval::launch_ref(const EvalArgs& a) {
    switch (a.head) {
        case BEAR_HEAD_REF_STOP: return launch_head<BEAR_HEAD_REF_STOP>(a);
        case BEAR_HEAD_REF_LINEAR: return launch_head<BEAR_HEAD_REF_LINEAR>(a);
    }
    return BEAR_ERR_ARG;
}
