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
