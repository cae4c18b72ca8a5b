"""the performance harness's output files (src/perf.hpp:364-461) with a B200 column."""
import re

import pytest

import perf_dat

pytestmark = pytest.mark.gpu


def test_dat_files_have_the_reference_format(tmp_path):
    files = perf_dat.run(tmp_path, ["int32-int32", "double-int64"], ["Uniform", "ZeroOne"], max_log2=12, seed=42, max_reps=4)
    names = sorted(f.name for f in files)
    assert "tpe-int32-int32-Uniform.dat" in names and "double-int64-ZeroOne-262144.dat" in names
    tpe = (tmp_path / "tpe-int32-int32-Uniform.dat").read_text().splitlines()
    hdr = tpe[0].split()
    assert hdr[0] == "number_of_elements" and hdr[1:3] == ["RadixB200", "RadixB200Host"]
    assert [int(r.split()[0]) for r in tpe[1:]] == [1 << i for i in range(13)]          # perf.hpp:391
    assert all(re.fullmatch(r"\d+( \d+\.\d{6})+", r) for r in tpe[1:])                   # std::fixed, setprecision(6)
    one = (tmp_path / "int32-int32-Uniform-262144.dat").read_text().split("\n")
    assert one[0] == "sort_method nanoseconds_per_element" and one[-2] == "" and one[1].startswith("RadixB200 ")
    ns = float(one[1].split()[1])
    assert 0 < ns < 50  # 2^18 int32 pairs: a few hundred microseconds at most
