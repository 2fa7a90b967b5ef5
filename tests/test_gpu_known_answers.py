"""The reference's closed-form known answers (SHAByteBufferTest.scala:225-331,534-700; ColumnUpdateDeleteTests.scala:205-332)
driven through the CUDA path: capi.Plan on cuda:0, compared with the closed forms themselves."""
import pytest

import known_answer_cases as K

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", K.CASES, ids=lambda c: c.__name__)
def test_known_answer_on_gpu(case, gpu_api):
    case(K.GpuEngine(gpu_api))
