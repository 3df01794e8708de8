"""The oracle against the committed golden fixtures (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py).
The fixtures freeze the oracle's answers (they are not outputs of the Rust reference, which cannot run here)."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_matches_committed_golden_fixtures():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    fresh = mg.build()
    gold = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
    assert sorted(fresh) == sorted(gold.files)
    for name in gold.files:
        assert np.array_equal(fresh[name], gold[name]), name


def test_golden_streams_have_the_reference_wire_layout():
    gold = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
    vor = gold["stream_voronoi6"].tobytes()
    assert len(vor) == 16 + 19 * 6  # clusterc.rs:155-165, 250-257
    assert vor[:4] == (24).to_bytes(4, "little") and vor[4:8] == (18).to_bytes(4, "little") and vor[8:16] == (6).to_bytes(8, "little")
    assert vor[16 + 8:16 + 16] == (3).to_bytes(8, "little")  # Rgb serialises as a slice: u64 length first (ser.rs:210-214)
    for name in ("stream_hufman", "stream_delta_stream", "stream_rle", "stream_ccol5"):
        s = gold[name].tobytes()
        assert s[:8] == (24).to_bytes(4, "little") + (18).to_bytes(4, "little")  # dims first (ser.rs:146-151)
    rle = gold["stream_rle"].tobytes()
    assert (len(rle) - 8) % 12 == 0  # records = u8 count + 11-byte Rgb (hilbertc.rs:35-38)
