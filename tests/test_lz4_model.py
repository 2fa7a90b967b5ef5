"""The device LZ4 decoder's ring / flush / dependency-round invariants, checked on the host with the executable model
of tools/lz4_model.py using the constants of snappydata_b200/csrc/sd_lz4.cu (the kernel itself is tested on the GPU
against liblz4 in tests/test_gpu_general.py::test_lz4_*)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_model_with_the_kernels_constants_decodes_adversarial_blocks():
    import lz4_model
    for which in ("LzDefault", "LzDense"):   # the measured shape and the denser, not yet measured one
        lz4_model.use(which)
        k = lz4_model.K
        assert k["LZ_WIN"] & (k["LZ_WIN"] - 1) == 0
        done = list(lz4_model.check(scale=4, alignments=(0, 5)))
        assert len(done) == 11
    # the window parse (kernel template parameter PARSE == 1): at every step it must return a prefix of what the serial
    # parse finds from the same position -- checked inside the model -- and the decode must stay byte-exact
    lz4_model.use("LzDense")
    assert len(list(lz4_model.check(scale=4, alignments=(8,), parse=1))) == 11
