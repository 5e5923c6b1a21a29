"""GPU tests (-m gpu) of the device-side round control (csrc/control.cu): the device sampler against the engine's host
sampler — which tests/test_ref_pins.py pins bit for bit to robust/sampling.cc compiled in place — for uniform and PROSAC
sampling, tiny point sets (most samples reject duplicates), state carried across rounds of every size."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cabi():
    from poselib_b200 import cabi as c
    if c.device_count() == 0:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box")
    c.set_device(0)
    return c


@pytest.mark.parametrize("n,k", [(10000, 5), (200, 3), (5000, 7), (20000, 4), (7, 7), (8, 7), (5, 5), (3, 3), (12, 4),
                                 (37, 5), (1 << 20, 7)])
@pytest.mark.parametrize("seed", [0, 1, 0xdeadbeefcafe])
def test_uniform_sampler_equals_host_sampler(cabi, n, k, seed):
    opt = cabi.RansacOpt(seed=seed)
    iters = 3000
    host = cabi.host_sample_table(n, k, opt, iters)
    for rnd in (1, 31, 32, 33, 1000, 4096):
        dev = cabi.device_sample_table(n, k, opt, iters, round_size=rnd)[0]
        assert np.array_equal(dev, host), (n, k, seed, rnd, np.argwhere((dev != host).any(axis=1))[:3])


def test_many_samplers_in_one_launch(cabi):
    iters, count = 2000, 37
    dev = cabi.device_sample_table(10000, 5, cabi.RansacOpt(seed=100), iters, round_size=512, count=count)
    for j in range(count):
        host = cabi.host_sample_table(10000, 5, cabi.RansacOpt(seed=100 + j), iters)
        assert np.array_equal(dev[j], host), j


@pytest.mark.parametrize("n,k,max_prosac", [(5000, 7, 100000), (500, 7, 3000), (200, 3, 1000), (64, 5, 200), (9, 7, 50),
                                            (7, 7, 10), (2000, 4, 1), (2000, 4, 0), (300, 5, 2)])
@pytest.mark.parametrize("seed", [0, 7])
def test_prosac_sampler_equals_host_sampler(cabi, n, k, max_prosac, seed):
    opt = cabi.RansacOpt(seed=seed, progressive_sampling=1, max_prosac_iterations=max_prosac)
    iters = 6000  # past max_prosac for most cases: the sampler falls back to uniform draws (sampling.cc:86,102)
    host = cabi.host_sample_table(n, k, opt, iters)
    for rnd in (1, 32, 257, 4096):
        dev = cabi.device_sample_table(n, k, opt, iters, round_size=rnd)[0]
        assert np.array_equal(dev, host), (n, k, max_prosac, seed, rnd, np.argwhere((dev != host).any(axis=1))[:3])
    os.environ["PLB_PROSAC_SEQUENTIAL"] = "1"  # the step-by-step subset growth instead of the closed form
    try:
        dev = cabi.device_sample_table(n, k, opt, iters, round_size=300)[0]
    finally:
        del os.environ["PLB_PROSAC_SEQUENTIAL"]
    assert np.array_equal(dev, host)


def test_c3_prosac_full_length(cabi):
    opt = cabi.RansacOpt(seed=0, progressive_sampling=1, max_prosac_iterations=100000)
    host = cabi.host_sample_table(5000, 7, opt, 100000)
    dev = cabi.device_sample_table(5000, 7, opt, 100000, round_size=16384)[0]
    assert np.array_equal(dev, host)
