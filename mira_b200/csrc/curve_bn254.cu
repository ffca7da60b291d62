// curve_bn254.cu — BN254 G1 instantiation: coordinates in Fq, scalars in Fr (y^2 = x^3 + 3).
#include "pipeline.cuh"
namespace mira_host {
using CF = mira::FqTag;
using SF = mira::FrTag;
static int check(mira_msm_ctx* c) { return check_on_curve_impl<CF>(c, 3u, 0); }
const CurveOps OPS_BN254 = {commit_impl<CF, SF>, commit_batch_impl<CF, SF>, prepare_impl<CF, SF>, check, combine_impl<CF>, partial_batch_dev_impl<CF, SF>, combine_dev_impl<CF>, partial_to_peer_impl<CF, SF>, gen_scalars_impl<SF>,
                            gen_bases_impl<CF, SF>, test_point_op_impl<CF>};
}  // namespace mira_host
