"""Layer-level parity of the tcgen05 kernel families (generic brick kernel incl. channel-chunked deep layers,
plane-sweep conv_ps, transposed plane-sweep conv_pst) against the CUDA-core kernels on identical bf16 CG8
inputs: <= 2e-2 of the output range (one bf16 ulp of the largest value), through the C ABI (sgm_debug_conv).
Groups "cs" (channel-streamed persistent kernel, conv_cs.cu) and "torch" additionally compare every case with torch's own
conv3d / conv_transpose3d on the same bf16 inputs and BN-folded bf16 weights (<= 1e-2 of the output range)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group", ["s1", "ps", "s2", "t2", "cs", "torch", "c20"])
def test_tc_layers_match_cuda_core_kernels(cuda_device, group):
    from tests import diag_tc_layers

    results = diag_tc_layers.run(group)
    assert results and all(results), f"{sum(results)}/{len(results)} cases passed in group {group}"


def test_row_sweep_head_matches_plane_sweep(cuda_device):
    """The row-sweep head kernel (conv_rs.cu: d0 and d1 taps folded into MMA N through overlapping accumulator
    columns) against the plane-sweep kernel (SGM_NO_RS=1) on identical inputs: whole-network forwards and
    sliding-window predictions, full / partial strips, 3 and 4 lane quarters, one and several units per CTA.
    Only the fp32 accumulation order differs: <= 2e-5 of the logit range."""
    from tests import diag_rs_head

    results = diag_rs_head.run()
    assert results and all(results), f"{sum(results)}/{len(results)} cases passed"
