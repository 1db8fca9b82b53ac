"""GPU tier: BASELINE.json's FULL sizes, checked through size-independent properties
(the oracle cannot finish these in seconds): linearity / sampled exactness for axpy,
"sum of partial sums" for reductions, difference-recovers-input for the scan, planted
extrema for argmax.  Data is generated on the device (seeded)."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


@pytest.fixture(scope='module')
def cp():
    import cupy_b200
    return cupy_b200


def gen(shape, dtype, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return (torch.rand(shape, device='cuda', generator=g, dtype=torch.float32) * 2 - 1).to(dtype)


def test_config2_axpy_2p28(cp):
    n = 1 << 28
    tx, ty = gen((n,), torch.float32, 1), gen((n,), torch.float32, 2)
    k = cp.ElementwiseKernel('T a, T x, T y', 'T z', 'z = a * x + y', 'axpy')
    z = k(np.float32(1.5), cp.from_torch(tx), cp.from_torch(ty))
    tz = z.to_torch()
    # bit-exact against the fmaf oracle on samples: head, tail, and a strided middle
    for sl in (slice(0, 1 << 16), slice(n - (1 << 16) - 3, n), slice(12345, n, 4099)):
        want = oracle.axpy(1.5, tx[sl].cpu().numpy(), ty[sl].cpu().numpy())
        np.testing.assert_array_equal(tz[sl].cpu().numpy(), want)
    # whole-array property: same bits as torch's fused addcmul-free fma form computed in float64 then rounded
    ref = (tx.double() * 1.5 + ty.double()).float()
    assert int((tz != ref).sum().item()) == 0        # double rounding cannot occur: 1.5*x is exact in float64, sum fits 53 bits
    del tz, z, ref


@pytest.mark.parametrize('tdt', [torch.float32, torch.float16])
def test_config3_axis_reductions_32768(cp, tdt):
    m = 32768
    t = gen((m, m), tdt, 5)
    # plant exact extrema: a unique max per row and per column position known in closed form
    rows = torch.arange(m, device='cuda')
    t[rows, (rows * 7 + 3) % m] = 2.0
    x = cp.from_torch(t)
    np_dt = np.float32 if tdt == torch.float32 else np.float16
    # sum: partial sums over blocks of rows add up (fp32 accumulate)
    s0, s1 = x.sum(axis=0), x.sum(axis=1)
    assert s0.dtype == np_dt and s1.dtype == np_dt and s0.shape == (m,)
    ref0 = t.double().sum(dim=0)
    ref1 = t.double().sum(dim=1)
    if tdt == torch.float32:
        tol = 1e-5 * t.double().abs().sum(dim=0).max().item()
        assert float((s0.to_torch().double() - ref0).abs().max().item()) <= tol
        assert float((s1.to_torch().double() - ref1).abs().max().item()) <= tol
    else:
        # fp32 accumulate, ONE rounding to fp16 at the end: the error is half an fp16 ulp of the result plus the
        # fp32 accumulation error (<= 1e-6 of sum|x|), not the 4-8 ulp a flat 0.5 would allow
        for got, ref, ax in ((s0, ref0, 0), (s1, ref1, 1)):
            ulp = torch.from_numpy(np.spacing(np.abs(ref.cpu().numpy()).astype(np.float16)).astype(np.float64)).cuda()
            bound = 0.5 * ulp + 1e-6 * t.double().abs().sum(dim=ax)
            assert bool(((got.to_torch().double() - ref).abs() <= bound).all().item())
    # max / argmax: the planted 2.0
    am1 = x.argmax(axis=1).to_torch()
    assert bool((am1 == (rows * 7 + 3) % m).all().item())
    assert bool((x.max(axis=1).to_torch() == 2.0).all().item())
    am0 = x.argmax(axis=0).to_torch()
    inv = torch.empty(m, dtype=torch.int64, device='cuda')
    inv[(rows * 7 + 3) % m] = rows                    # 7 is coprime to 2^15: a permutation
    assert bool((am0 == inv).all().item())
    assert bool((x.max(axis=0).to_torch() == 2.0).all().item())
    # var: against float64 two-pass on the device
    for ax in (0, 1):
        v = x.var(axis=ax).to_torch().double()
        ref = t.double().var(dim=ax, unbiased=False)
        rel = ((v - ref).abs() / ref).max().item()
        assert rel <= (1e-5 if tdt == torch.float32 else 2.0 ** -10), rel      # fp16: one ulp is 2^-10 relative
    del t


def test_config4a_transposed_exp_broadcast(cp):
    base = gen((256, 1024, 1024), torch.float32, 9)
    v = gen((256,), torch.float32, 10)
    xt = cp.from_torch(base).transpose(2, 1, 0)
    assert xt.shape == (1024, 1024, 256) and xt.strides == (4, 4096, 4194304)
    fused = cp.ElementwiseKernel('T x, T v', 'T z', 'z = exp(x) + v', 'exp_add')
    z1 = fused(xt, cp.from_torch(v))
    z2 = cp.exp(xt) + cp.from_torch(v)
    assert z1.flags.c_contiguous and z1.shape == (1024, 1024, 256)
    ref = (torch.exp(base.double()).permute(2, 1, 0) + v.double())
    for z in (z1, z2):
        err = (z.to_torch().double() - ref).abs()
        bound = 2 * 2.0 ** -23 * torch.exp(base.double()).permute(2, 1, 0) + 2.0 ** -23 * ref.abs()
        assert bool((err <= bound).all().item())
    assert bool((z1.to_torch() == z2.to_torch()).all().item())      # same arithmetic, one launch vs two
    del base, ref


def test_config4b_cumsum_int64_2p28(cp):
    n = 1 << 28
    g = torch.Generator(device='cuda').manual_seed(11)
    t = torch.randint(-(1 << 20), 1 << 20, (n,), device='cuda', generator=g, dtype=torch.int64)
    y = cp.cumsum(cp.from_torch(t)).to_torch()
    assert y.dtype == torch.int64
    # differences recover the input; the last element is the total; samples agree with the oracle
    assert bool((y[1:] - y[:-1] == t[1:]).all().item()) and int(y[0].item()) == int(t[0].item())
    assert int(y[-1].item()) == int(t.sum().item())
    np.testing.assert_array_equal(y[:100000].cpu().numpy(), oracle.cumsum(t[:100000].cpu().numpy()))
    del y, t


def test_casting_scans_2p28(cp):
    """the casting flat scans at full size, one read of x, bit-exact for integers.
    Size-independent properties: differences recover the input, the last element is the total; the ragged end
    (n not a multiple of the 128-item row granule) goes through the tail path."""
    n = (1 << 28) + 77
    g = torch.Generator(device='cuda').manual_seed(12)
    t32 = torch.randint(-1000, 1000, (n,), device='cuda', generator=g, dtype=torch.int32)
    y = cp.cumsum(cp.from_torch(t32)).to_torch()
    assert y.dtype == torch.int64
    assert bool((y[1:] - y[:-1] == t32[1:]).all().item()) and int(y[0]) == int(t32[0])
    assert int(y[-1]) == int(t32.sum(dtype=torch.int64))
    np.testing.assert_array_equal(y[-100000:].cpu().numpy() - int(y[-100001]),
                                  oracle.cumsum(t32[-100000:].cpu().numpy().astype(np.int64)))
    del y, t32
    tb = torch.rand(n, device='cuda', generator=g) < 0.3
    y = cp.cumsum(cp.from_torch(tb)).to_torch()
    assert y.dtype == torch.int64
    assert bool((y[1:] - y[:-1] == tb[1:].to(torch.int64)).all().item()) and int(y[-1]) == int(tb.sum())
    del y, tb
    # float16 with a float accumulator: one rounding per output (half an fp16 ulp) + the fp32 accumulation error
    th = ((torch.rand(n, device='cuda', generator=g) * 2 - 1) / 64).half()
    y = cp.cumsum(cp.from_torch(th)).to_torch()
    assert y.dtype == torch.float16
    ref = torch.cumsum(th.double(), 0)
    ulp = torch.from_numpy(np.spacing(np.abs(ref.cpu().numpy()).astype(np.float16)).astype(np.float64)).cuda()
    bound = 0.5 * ulp + 1e-6 * torch.cumsum(th.double().abs(), 0)
    assert bool(((y.double() - ref).abs() <= bound).all().item())
    del y, th, ref, ulp, bound


def test_cumsum_axis0_16384_squared(cp):
    """Column-march scan at the size DESIGN.md quotes: float32 within tolerance, int64 bit-exact."""
    m = 16384
    t = gen((m, m), torch.float32, 13)
    y = cp.cumsum(cp.from_torch(t), axis=0).to_torch()
    ref_last = t.double().sum(dim=0)
    assert float((y[-1].double() - ref_last).abs().max()) <= 1e-5 * float(t.double().abs().sum(dim=0).max())
    assert bool(torch.allclose(y[1:] - y[:-1], t[1:], atol=2e-3))           # differences recover the input
    np.testing.assert_allclose(y[:, :64].cpu().numpy(), np.cumsum(t[:, :64].cpu().numpy().astype(np.float64), axis=0),
                               atol=2e-3)
    del y
    ti = torch.randint(-1000, 1000, (m // 2, m), device='cuda', dtype=torch.int64)
    yi = cp.cumsum(cp.from_torch(ti), axis=0).to_torch()
    assert bool(torch.equal(yi, torch.cumsum(ti, 0)))
    del t, ti, yi


def test_full_reductions_2p28_and_beyond_int32_indexing(cp):
    n = 1 << 28
    t = gen((n,), torch.float32, 21)
    x = cp.from_torch(t)
    ref = t.double().sum().item()
    assert abs(float(x.sum().get()) - ref) <= 1e-5 * t.double().abs().sum().item()
    t[123456789] = 5.0
    assert int(x.argmax().get()) == 123456789 and float(x.max().get()) == 5.0
    v = float(x.var().get())
    assert abs(v - t.double().var(unbiased=False).item()) <= 1e-5 * v
    del t, x
    # > 2^31 elements (64-bit indexing; tests/cupy_tests/core_tests/test_reduction.py:86-93,
    # sorting_tests/test_search.py:88-93)
    big = torch.ones((1 << 31) + 5, dtype=torch.int8, device='cuda')
    xb = cp.from_torch(big)
    assert int(xb.sum().get()) == (1 << 31) + 5
    big[(1 << 31) + 2] = 3
    assert int(xb.argmax().get()) == (1 << 31) + 2
    assert int((xb + xb).sum().get()) == 2 * ((1 << 31) + 5) + 4
    del big, xb
