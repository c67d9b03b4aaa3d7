"""Layer-level parity of the tcgen05 kernel families (generic brick kernel incl. channel-chunked deep layers,
plane-sweep conv_ps, transposed plane-sweep conv_pst) against the CUDA-core kernels on identical bf16 CG8
inputs: <= 2e-2 of the output range (one bf16 ulp of the largest value), through the C ABI (sgm_debug_conv)."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group", ["s1", "ps", "s2", "t2"])
def test_tc_layers_match_cuda_core_kernels(cuda_device, group):
    from tests import diag_tc_layers

    results = diag_tc_layers.run(group)
    assert results and all(results), f"{sum(results)}/{len(results)} cases passed in group {group}"
