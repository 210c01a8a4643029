"""Host-side data plumbing against outputs of the REAL reference functions (tests/golden/make_data_golden.py):
the agent partitioning of main.py:524-682, the train/test split of main.py:2353-2361 and the SRTM .hgt loader of
real_world_datasets.py:238-572.  Index and byte work: every comparison is exact."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

import dqgp_b200 as d

_spec = importlib.util.spec_from_file_location("make_data_golden", os.path.join(GOLDEN, "make_data_golden.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)
CASES = G.cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_split_data_numpy_selects_the_reference_rows(name):
    gold = load_golden("data_plumbing.npz")
    x, y, n_agents, method, pct, seed = CASES[name]
    shards = d.split_data_numpy(x, y, n_agents, method, pct, seed)
    assert len(shards) == n_agents
    assert np.array_equal([len(ya) for _, ya in shards], gold[f"split_{name}_sizes"])
    assert np.array_equal(np.concatenate([ya for _, ya in shards]), gold[f"split_{name}_rows"])
    for xa, ya in shards:
        assert np.array_equal(xa, x[(ya - 0.5).astype(int)])
    if pct == 1.0:
        idx = d.split_indices(x, n_agents, method, seed)
        assert np.array_equal(np.concatenate(idx) + 0.5, gold[f"split_{name}_rows"])


def test_points_on_interior_grid_faces_go_to_both_neighbours():
    """main.py:580-583 tests both faces inclusively: the reference duplicates those rows, and so does the drop-in."""
    x, y, n_agents, method, _, seed = CASES["grid_2d_4_points_on_faces"]
    rows = np.concatenate(d.split_indices(x, n_agents, method, seed))
    assert len(rows) == 100 and len(np.unique(rows)) == 81


def test_unknown_partition_and_bad_percentage_raise_like_the_reference():
    x, y = CASES["kd_2d_3"][:2]
    with pytest.raises(ValueError, match="Unknown partition method"):
        d.split_data_numpy(x, y, 3, "spiral")
    for pct in (0.0, -0.1, 1.5):
        with pytest.raises(ValueError, match="Percentage must be between"):
            d.sample_agent_data_percentage(x, y, pct)


def test_prepare_training_data_matches_main():
    gold = load_golden("data_plumbing.npz")
    x, rows = CASES["kd_2d_8"][:2]                              # labels: row number + 0.5
    shards, (x_tr, r_tr), (x_te, r_te) = d.prepare_training_data(x, rows, 8, "regional", test_split=0.1, seed=42, split_seed=43)
    assert np.array_equal(r_tr, gold["tts_train_rows"]) and np.array_equal(r_te, gold["tts_test_rows"])
    assert np.array_equal(x_tr, x[(r_tr - 0.5).astype(int)]) and np.array_equal(x_te, x[(r_te - 0.5).astype(int)])
    assert np.array_equal([len(r) for _, r in shards], gold["tts_shard_sizes"])
    assert np.array_equal(np.concatenate([r for _, r in shards]), gold["tts_shard_rows"])


@pytest.fixture(scope="module")
def tile_dir(tmp_path_factory):
    root = tmp_path_factory.mktemp("srtm")
    G.write_tiles(str(root))
    return os.path.join(str(root), "srtm_data")


def test_read_hgt_file(tile_dir, tmp_path):
    gold = load_golden("data_plumbing.npz")
    tile = d.read_hgt_file(os.path.join(tile_dir, "N17E073.hgt"))
    assert tile.shape == (1201, 1201) and tile.dtype == np.float64
    assert np.array_equal(tile[::100, ::100], gold["hgt_tile"])
    assert np.array_equal(tile, G.synthetic_tile().astype(np.float64)) and (tile == -32768).sum() > 1000
    bad = tmp_path / "short.hgt"
    bad.write_bytes(b"\0" * 1000)
    with pytest.raises(ValueError, match="Unexpected HGT file size: 1000 bytes"):
        d.read_hgt_file(str(bad))
    big = tmp_path / "one_arc_second.hgt"                      # 3601 x 3601, big-endian
    np.arange(3601 * 3601, dtype=np.int64).astype(">i2").tofile(str(big))
    one = d.read_hgt_file(str(big))
    assert one.shape == (3601, 3601) and one[0, 258] == 258.0 and one[1, 0] == 3601.0


@pytest.mark.parametrize("name", sorted(G.srtm_cases()))
def test_load_srtm_elevation_dataset_equals_reference(tile_dir, name):
    gold = load_golden("data_plumbing.npz")
    x, y = d.load_srtm_elevation_dataset(data_dir=tile_dir, **G.srtm_cases()[name])
    assert x.dtype == np.float64 and y.dtype == np.float64
    assert np.array_equal(x, gold[f"srtm_{name}_X"]) and np.array_equal(y, gold[f"srtm_{name}_Y"])


def test_config2_pipeline_rebuilds_the_shards_of_the_recorded_main_run(tile_dir):
    """BASELINE configs[1] end to end on the host side: tile -> loader (main.py:2145-2160) -> train/test split with the data
    seed (main.py:2355) -> regional partition (main.py:2372) gives exactly the training set and the four shards the real
    main.main() handed to its agents in tests/golden/trajectory_cfg2_srtm.json (the device replays that run in
    tests/test_gpu_agent.py)."""
    import json
    with open(os.path.join(GOLDEN, "trajectory_cfg2_srtm.json")) as f:
        rec = json.load(f)
    data = load_golden("trajectory_cfg2_srtm_data.npz")
    x, y = d.load_srtm_elevation_dataset(region="maharashtra", max_samples=1000, subsample_factor=10, normalize=True,
                                         random_state=rec["srtm_data_seed"], data_dir=tile_dir)
    shards, (x_tr, y_tr), (x_te, y_te) = d.prepare_training_data(x, y, rec["n_agents"], "regional", test_split=0.1, seed=42,
                                                                 split_seed=rec["srtm_data_seed"])
    assert len(y_te) == 100 and np.array_equal(x_tr, data["X_train"]) and np.array_equal(y_tr, data["Y_train"])
    assert len(shards) == 4
    for a, (xa, ya) in enumerate(shards):
        assert np.array_equal(xa, data[f"X_{a}"]) and np.array_equal(ya, data[f"Y_{a}"])
    assert sum(len(ya) for _, ya in shards) > len(y_tr)          # normalised grid coordinates: rows on the cell faces are shared


def test_srtm_loader_errors(tile_dir):
    with pytest.raises(ValueError, match="not supported"):
        d.load_srtm_elevation_dataset(region="atlantis", data_dir=tile_dir)
    with pytest.raises(FileNotFoundError, match="HGT file not found for tile N47W124"):
        d.load_srtm_elevation_dataset(region="washington_coast", data_dir=tile_dir)
    with pytest.raises(FileNotFoundError, match="Preprocessed file not found"):
        d.load_srtm_elevation_dataset(region="maharashtra", use_preprocessed=True, preprocessed_dir=tile_dir)
    assert d.get_tile_for_region("oregon_coast") == "N45W123" and d.get_tile_for_region("N00E000") == "N00E000"


@pytest.mark.parametrize("dim,n,noise,seed", [(1, 57, 0.1, 3), (2, 64, 0.05, 11), (3, 41, 0.2, 2024)])
def test_generate_data_numpy_equals_reference(dim, n, noise, seed):
    """The --classical-dataset generator (main.py:457-522): the inputs bit for bit (same random stream), the targets to rounding."""
    gold = load_golden("data_plumbing.npz")
    x, y = d.generate_data_numpy(n, dim, noise, seed)
    assert np.array_equal(x, gold[f"classical_{dim}d_X"])
    assert np.max(np.abs(y - gold[f"classical_{dim}d_Y"])) < 1e-13
    with pytest.raises(ValueError, match="Unsupported input dimension: 4"):
        d.generate_data_numpy(5, 4, 0.1, 1)


@pytest.mark.parametrize("tag", ["novar", "var"])
def test_evaluate_predictions_equals_reference(tag):
    """The metric dictionary of main.py:1598-1736 (scikit-learn's mse / mae / r2 there, plain NumPy here): same keys, numbers to
    1e-12, same verbal grades; three predicted variances are zero (NLPD floor 1e-10, infinite calibration rmse as in the reference)."""
    gold = load_golden("data_plumbing.npz")
    y_true, y_pred, y_var = G.prediction_case()
    with np.errstate(divide="ignore", invalid="ignore"):
        got = d.evaluate_predictions(y_true, y_pred, None if tag == "novar" else y_var)
    want = {k[len(f"metrics_{tag}_"):]: gold[k] for k in gold.files if k.startswith(f"metrics_{tag}_")}
    assert set(got) == set(want)
    for k, v in want.items():
        if v.dtype.kind in "US":
            assert got[k] == str(v), k
        elif np.isfinite(v):
            assert abs(got[k] - float(v)) <= 1e-12 * max(1.0, abs(float(v))), k
        else:
            assert not np.isfinite(got[k]) and np.isnan(got[k]) == np.isnan(v), k


def test_save_quantum_dataset_writes_the_reference_csv(tmp_path):
    gold = load_golden("data_plumbing.npz")
    x, y = CASES["grid_2d_4"][:2]
    path = d.save_quantum_dataset(x[:5], y[:5], "golden", output_dir=str(tmp_path / "out"))
    assert os.path.basename(path) == str(gold["csv_name"]) == "golden_2d_5.csv"
    assert open(path).read() == str(gold["csv_text"])
