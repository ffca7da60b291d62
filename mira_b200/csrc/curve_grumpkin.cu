// curve_grumpkin.cu — Grumpkin G1 instantiation: coordinates in Fr, scalars in Fq (y^2 = x^3 - 17).
#include "pipeline.cuh"
namespace mira_host {
using CF = mira::FrTag;
using SF = mira::FqTag;
static int check(mira_msm_ctx* c) { return check_on_curve_impl<CF>(c, 17u, 1); }
const CurveOps OPS_GRUMPKIN = {commit_impl<CF, SF>, commit_batch_impl<CF, SF>, prepare_impl<CF, SF>, check, combine_impl<CF>, partial_batch_dev_impl<CF, SF>, combine_dev_impl<CF>, partial_to_peer_impl<CF, SF>, gen_scalars_impl<SF>,
                               gen_bases_impl<CF, SF>, test_point_op_impl<CF>};
}  // namespace mira_host
