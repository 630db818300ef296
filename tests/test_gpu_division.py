"""The exact kernels divide by a shared reciprocal (csrc/common.cuh: make_divisor / fast_quotient / xdiv).
The quotients must be the IEEE-rounded ones the reference's C simulation computes with `/`
(e.g. LU/lu_v1.1/model4x4/lu.cpp:70, Cholesky/cholesky_v1.3/model4x4/chol.cpp:47): checked bit for bit
on the device against __fdiv_rn, and against numpy's float32 division on the host."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _pairs(seed, count):
    g = torch.Generator(device="cuda").manual_seed(seed)
    # random bit patterns restricted to finite values: every mantissa, every exponent
    a = torch.randint(-(2 ** 31), 2 ** 31 - 1, (count,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32).view(torch.float32)
    b = torch.randint(-(2 ** 31), 2 ** 31 - 1, (count,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32).view(torch.float32)
    return a, b


def test_all_bit_patterns_sample():
    import hls_designs_b200 as hd
    a, b = _pairs(1, 1 << 26)
    bad, fast, _ = hd.selftest_division(a, b)
    assert bad == 0
    assert fast > 0  # a minority of random exponents falls in the fast range


def test_moderate_exponents_take_the_fast_path():
    import hls_designs_b200 as hd
    g = torch.Generator(device="cuda").manual_seed(2)
    count = 1 << 27
    # mantissas uniform, exponents within 2^-40 .. 2^40: the range elimination values live in
    def rnd():
        m = torch.randint(0, 1 << 23, (count,), generator=g, device="cuda", dtype=torch.int32)
        e = torch.randint(127 - 40, 127 + 40, (count,), generator=g, device="cuda", dtype=torch.int32)
        s = torch.randint(0, 2, (count,), generator=g, device="cuda", dtype=torch.int32)
        return ((s << 31) | (e << 23) | m).view(torch.float32)
    a, b = rnd(), rnd()
    bad, fast, _ = hd.selftest_division(a, b)
    # (divisors whose mantissa is all ones - about 2^-23 of the draws - are routed to IEEE division)
    assert bad == 0 and count - 256 <= fast <= count


def test_special_values_and_host_agreement():
    import hls_designs_b200 as hd
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1.17549435e-38, 3.4028235e38,
                   2.0 ** -60, 2.0 ** 60, np.nextafter(np.float32(2.0 ** -60), np.float32(0)),
                   np.nextafter(np.float32(2.0 ** 60), np.float32(np.inf)), 3.0, 1.0 / 3.0, 100.0, 7.0], dtype=np.float32)
    a, b = np.meshgrid(sp, sp, indexing="ij")
    a, b = a.ravel().copy(), b.ravel().copy()
    bad, _, _ = hd.selftest_division(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    assert bad == 0
    # integer-valued operands as the testbenches generate them (values 1..100): a/b vs numpy float32
    rng = np.random.default_rng(3)
    x = rng.integers(1, 101, 1 << 20).astype(np.float32)
    y = rng.integers(1, 101, 1 << 20).astype(np.float32)
    bad, fast, _ = hd.selftest_division(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())
    assert bad == 0 and fast == x.size


def test_hard_divisors_all_ones_mantissa():
    """Round-1 advisor finding: with a reciprocal that is one ulp low, q0 + r*rem lands on a false tie for divisors
    whose mantissa is all ones and power-of-two numerators (0x1p23 / 0x1.fffffep23 must give 0x1.000002p-1).
    Every exponent of the fast range x divisor mantissas {all ones, all ones - 1, 0, 1, 0x400000} x numerators
    {powers of two, all-ones, near-midpoint patterns, random}: bit-identical to IEEE division."""
    import hls_designs_b200 as hd
    exps = np.arange(127 - 60, 127 + 61, dtype=np.uint32)
    dm = np.array([0x7FFFFF, 0x7FFFFE, 0x000000, 0x000001, 0x400000, 0x3FFFFF], dtype=np.uint32)
    rng = np.random.default_rng(11)
    nm = np.concatenate([np.array([0x000000, 0x7FFFFF, 0x7FFFFE, 0x000001, 0x400000, 0x400001, 0x3FFFFF], dtype=np.uint32),
                         rng.integers(0, 1 << 23, 57, dtype=np.uint32)])
    den = ((exps[:, None] << 23) | dm[None, :]).ravel()
    num = ((exps[:, None] << 23) | nm[None, :]).ravel()
    a, b = np.meshgrid(num, den, indexing="ij")
    a = np.ascontiguousarray(a.ravel()).view(np.float32)
    b = np.ascontiguousarray(b.ravel()).view(np.float32)
    b[::2] *= -1.0
    bad, fast, unguarded = hd.selftest_division(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    print(f"{a.size} pairs: {bad} mismatches, {fast} on the fast sequence, {unguarded} would differ without the guard")
    assert bad == 0
    assert 0 < fast < a.size  # all-ones divisors take IEEE division
    want = a / b  # numpy float32 division is IEEE
    t = torch.from_numpy(a).cuda() / torch.from_numpy(b).cuda()
    assert np.array_equal(t.cpu().numpy().view(np.uint32), want.view(np.uint32))
