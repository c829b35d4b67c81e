"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/fh_b200.h declares; the product refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re
import pytest
import torch
from conftest import ROOT


def _declared(header="fh_b200.h"):
	src = open(os.path.join(ROOT, "include", header)).read()
	src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
	return sorted(set(re.findall(r"\b(fh_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
	import __graft_entry__ as ge
	ge.build()
	lib = ctypes.CDLL(ge.LIB)
	names = _declared()
	assert len(names) >= 15
	for n in names:
		assert hasattr(lib, n), "missing export %s" % n
	from fasthigashi_b200 import _lib
	assert sorted(_lib.EXPORTS) == names


def test_host_library_exports_every_declared_symbol():
	import __graft_entry__ as ge
	ge.build_host()
	lib = ctypes.CDLL(ge.HOST_LIB)
	names = _declared("fh_host.h")
	assert len(names) >= 6
	for n in names:
		assert hasattr(lib, n), "missing export %s" % n
	from fasthigashi_b200 import ingest
	assert sorted(ingest.EXPORTS) == names
	assert ingest.lib().fh_host_version() >= 100


def test_version_and_error_string():
	from fasthigashi_b200 import _lib
	L = _lib.lib()
	assert L.fh_version() >= 100
	assert isinstance(L.fh_last_error(), bytes)


def test_no_cpu_fallback():
	from fasthigashi_b200 import _lib
	from fasthigashi_b200.partial_rwr import partial_rwr
	from fasthigashi_b200.project2orthogonal import project2orthogonal
	from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
	with pytest.raises(_lib.FHError):
		partial_rwr(torch.ones(2, 4, 6), 0, 4, True, True, False)
	with pytest.raises(_lib.FHError):
		project2orthogonal(torch.ones(2, 6, 3), 3, None)
	with pytest.raises(_lib.FHError):
		Fast_Higashi_core(8, 10, [1000000]).to("cpu")


def test_product_never_imports_oracle():
	pkg = os.path.join(ROOT, "fast-higashi_b200")
	for f in os.listdir(pkg):
		if f.endswith(".py"):
			src = open(os.path.join(pkg, f)).read()
			assert "oracle" not in src.replace("no oracle", ""), f
			assert "/root/reference" not in src, f


def test_argument_validation_returns_codes_before_any_launch():
	"""Error convention of the C ABI (include/fh_b200.h): bad descriptors come back as a non-zero status with the reason in
	fh_last_error(), decided on the host before any CUDA call - so it can be checked without a GPU."""
	import ctypes as C
	from fasthigashi_b200 import _lib
	L = _lib.lib()
	d = _lib.GemmDesc()
	d.M = d.N = d.K = 4; d.batch = 1; d.sa_m, d.sa_k, d.sb_k, d.sb_n, d.ldc, d.alpha = 2, 2, 4, 1, 4, 1.0
	assert L.fh_gemm_batched(C.byref(d), 16, 16, 16, None) != 0 and b"unit stride" in L.fh_last_error()
	d.sa_m, d.sa_k, d.dtype = 4, 1, 99
	assert L.fh_gemm_batched(C.byref(d), 16, 16, 16, None) != 0 and b"dtype" in L.fh_last_error()
	n = C.c_int(0)
	r = _lib.rwr_desc(16, 30, 30, 0, 3, True, True, False, 0, 4, 0)             # ldw not a multiple of 4
	assert L.fh_rwr_batched(C.byref(r), 16, 16, 16, None, 0, 16, 480, 16, 1 << 30, C.byref(n), None) != 0
	assert b"ldw" in L.fh_last_error()
	r = _lib.rwr_desc(16, 30, 32, 20, 3, True, True, False, 0, 4, 0)            # diagonal block sticks out of the window
	assert L.fh_rwr_batched(C.byref(r), 16, 16, 16, None, 0, 16, 512, 16, 1 << 30, C.byref(n), None) != 0
	assert b"outside window" in L.fh_last_error()
	r = _lib.rwr_desc(115, 315, 316, 100, 4, True, True, False, 0, 64, 0)
	need = L.fh_rwr_workspace_bytes(C.byref(r))
	assert need > 0
	assert L.fh_polar_batched(16, 16, 1, 8, 4, 4, 32, None, None, 0, 16, 0, None, None) != 0 and b"workspace" in L.fh_last_error()
	assert L.fh_polar_workspace_bytes(1, 8, 4) > 0
	assert L.fh_cp_als(16, 4, 2, 3, 16, 16, 16, 1, 16, 0, None, None) != 0 and b"workspace" in L.fh_last_error()
	assert L.fh_cp_als_workspace_bytes(4, 2, 3) > 0
	with pytest.raises(_lib.FHError, match="unit stride"):
		_lib.check(L.fh_gemm_batched(C.byref(_bad_gemm()), 16, 16, 16, None))


def _bad_gemm():
	from fasthigashi_b200 import _lib
	d = _lib.GemmDesc()
	d.M = d.N = d.K = 4; d.batch = 1; d.sa_m, d.sa_k, d.sb_k, d.sb_n, d.ldc, d.alpha = 2, 2, 4, 1, 4, 1.0
	return d


def _prototypes(header="fh_b200.h"):
	"""{name: [parameter declarations]} of every `fh_*` function the header declares (comments stripped)."""
	src = open(os.path.join(ROOT, "include", header)).read()
	src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
	src = re.sub(r"//[^\n]*", "", src)
	out = {}
	for m in re.finditer(r"\b(fh_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", src, flags=re.S):
		params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
		out[m.group(1)] = [] if params in ([""], ["void"]) else params
	return out


def _c_class(decl):
	if "*" in decl:
		return "ptr"
	for prefix, cls in (("long long", "i64"), ("int64_t", "i64"), ("size_t", "size"), ("double", "f64"), ("int32_t", "int"), ("int", "int")):
		if decl.startswith(prefix):
			return cls
	raise AssertionError("unclassified parameter: " + decl)


def _py_class(t):
	import ctypes as C
	if t in (C.c_void_p, C.c_char_p) or hasattr(t, "_type_") and not isinstance(t._type_, str):
		return "ptr"
	return {C.c_longlong: "i64", C.c_int64: "i64", C.c_size_t: "size", C.c_double: "f64", C.c_int: "int", C.c_int32: "int"}[t]


def test_host_ctypes_prototypes_agree_with_the_header():
	"""The same check for libfh_host.so (include/fh_host.h against ingest.py)."""
	from fasthigashi_b200 import ingest
	L = ingest.lib()
	protos = _prototypes("fh_host.h")
	assert set(protos) == set(ingest.EXPORTS)
	checked = 0
	for name, params in protos.items():
		at = getattr(getattr(L, name), "argtypes", None)
		if at is None:
			assert not params, name
			continue
		assert len(at) == len(params), (name, len(at), params)
		for t, decl in zip(at, params):
			assert _py_class(t) == _c_class(decl), (name, decl, t)
		checked += 1
	assert checked >= 6


def test_ctypes_prototypes_agree_with_the_header():
	"""The Python binding (`_lib.py`: argtypes) against the C prototypes of include/fh_b200.h: same argument count for every entry
	point that has argtypes, and the same coarse class per argument (pointer / 64-bit integer / int / size_t / double) - an edit
	of one side without the other would pass garbage through ctypes silently."""
	import ctypes as C
	from fasthigashi_b200 import _lib
	L = _lib.lib()
	protos = _prototypes()
	assert set(protos) == set(_lib.EXPORTS)

	checked = 0
	for name, params in protos.items():
		at = getattr(getattr(L, name), "argtypes", None)
		if at is None:
			assert not params or name in ("fh_rwr_workspace_bytes",), name    # parameterless getters need no argtypes
			continue
		assert len(at) == len(params), (name, len(at), params)
		for t, decl in zip(at, params):
			assert _py_class(t) == _c_class(decl), (name, decl, t)
		checked += 1
	assert checked >= 18
