"""CPU restatement of the operand format of the 3xFP16 RWR kernel (fast-higashi_b200/csrc/fh_rwr_chain16.cu, fh_rwr.cu,
fh_tc.cuh): the power-of-two scale taken from the exponent bits of a cell's largest value, the hi / lo binary16 split, the
products the tensor core forms from them. Pins the numbers DESIGN.md 3.1 states (22 significand bits above 2^-14 of the
largest value, an absolute resolution below, ~1e-3 relative at the 1e-8 floor, operand error of a chain at the 1e-7 level)
without a GPU. numpy's float16 conversion rounds to nearest even, like cvt.rn.f16x2.f32."""
import numpy as np


def scale_from_amax(amax):
	"""sa = 2^(13 - e) with e the unbiased exponent of max(amax, 1e-8): (267 - biased_exponent) << 23, as in the kernels."""
	a = np.maximum(np.float32(amax), np.float32(1e-8))
	bits = a.view(np.uint32) if isinstance(a, np.ndarray) else np.array([a], dtype=np.float32).view(np.uint32)
	return ((np.uint32(267) - (bits >> np.uint32(23))) << np.uint32(23)).view(np.float32)


def split(x):
	"""hi = rn16(x), lo = rn16(x - hi) for pre-scaled fp32 x (f16_split2)."""
	x = np.asarray(x, dtype=np.float32)
	hi = x.astype(np.float16)
	lo = (x - hi.astype(np.float32)).astype(np.float16)
	return hi, lo


def test_scale_puts_the_largest_value_in_2_13_2_14():
	rng = np.random.default_rng(0)
	amax = np.concatenate([10.0 ** rng.uniform(-8, 6, 2000), [1e-8, 1.0, 2.0, 8191.999, 8192.0, 0.0, 1e-12]]).astype(np.float32)
	sa = scale_from_amax(amax)
	scaled = np.maximum(amax, np.float32(1e-8)) * sa
	assert np.all(scaled >= 2.0 ** 13) and np.all(scaled < 2.0 ** 14)
	# exact powers of two: scaling is exact, and so is undoing it with (biased_exponent - 13) << 23
	m, _ = np.frexp(sa)
	assert np.all(m == 0.5)
	bits = np.maximum(amax, np.float32(1e-8)).view(np.uint32)
	sa_inv = (((bits >> np.uint32(23)) - np.uint32(13)) << np.uint32(23)).view(np.float32)
	assert np.all(sa * sa_inv == 1.0)


def test_split_carries_22_bits_above_2_to_minus_14_of_the_largest_value():
	rng = np.random.default_rng(1)
	amax = np.float32(5.0)
	sa = scale_from_amax(amax)[0]
	# scaled values >= 1/2 (the largest one is in [2^13, 2^14): "2^-14 of the largest value", to within its binade)
	x = (10.0 ** rng.uniform(np.log10(0.5 / float(sa)), np.log10(5.0), 200000)).astype(np.float32)
	hi, lo = split(x * sa)
	rec = (hi.astype(np.float64) + lo.astype(np.float64)) / float(sa)
	rel = np.abs(rec - x.astype(np.float64)) / x
	assert rel.max() <= 2.0 ** -21  # hi to 2^-11 of x, lo to 2^-11 of the remainder: 2^-22 (+ ties)
	assert np.sqrt(np.mean(rel ** 2)) < 1.5e-7


def test_small_entries_keep_an_absolute_resolution_and_the_floor_its_three_digits():
	amax = np.float32(5.0)
	sa = scale_from_amax(amax)[0]  # 2^11
	assert sa == 2048.0
	rng = np.random.default_rng(2)
	x = (10.0 ** rng.uniform(-8, np.log10(0.4999 / float(sa)), 100000)).astype(np.float32)  # scaled values below 1/2
	hi, lo = split(x * sa)
	rec = (hi.astype(np.float64) + lo.astype(np.float64)) / float(sa)
	assert np.abs(rec - x).max() <= 2.0 ** -25 / float(sa) * 1.0001  # half a binary16 subnormal step of the scaled value
	floor = np.float32(1e-8)
	h, l = split(np.array([floor * sa]))
	got = (float(h[0]) + float(l[0])) / float(sa)
	assert abs(got - 1e-8) / 1e-8 < 2e-3  # the tolerance of test_full_size_rwr_conserves_column_mass
	# a cell without contacts scales its floor into the normal range: exact to 22 bits
	sa0 = scale_from_amax(np.float32(0.0))[0]
	h, l = split(np.array([floor * sa0]))
	assert abs((float(h[0]) + float(l[0])) / float(sa0) - 1e-8) / 1e-8 < 2.0 ** -21


def test_three_products_reproduce_an_fp32_chain_step():
	"""hi hi + hi lo + lo hi with exact products and fp64 accumulation (only the operand representation is measured):
	one step Q <- 1/2 Q P + 1/2 I on a column-stochastic P with entries down to 1e-9, Q and P scaled by 2^14."""
	rng = np.random.default_rng(3)
	n = 115
	P = rng.random((n, n)) ** 6 + 1e-9
	P /= P.sum(0, keepdims=True)
	Q = 0.5 * P + 0.5 * np.eye(n)
	qs = 16384.0
	qh, ql = split(Q * qs)
	ph, pl = split(P * qs)
	f = lambda a: a.astype(np.float64)
	acc = f(qh) @ f(ph) + f(qh) @ f(pl) + f(ql) @ f(ph)
	got = 0.5 * acc / qs ** 2 + 0.5 * np.eye(n)
	ref = 0.5 * Q.astype(np.float32).astype(np.float64) @ P.astype(np.float32).astype(np.float64) + 0.5 * np.eye(n)
	err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
	assert err < 2e-7, err
	# one product (hi hi) alone is two orders worse than the tolerance allows: why the split is needed
	one = 0.5 * (f(qh) @ f(ph)) / qs ** 2 + 0.5 * np.eye(n)
	assert np.linalg.norm(one - ref) / np.linalg.norm(ref) > 1e-5
