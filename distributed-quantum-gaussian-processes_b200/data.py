"""Host-side data plumbing either side of the hot path: the SRTM ``.hgt`` tile reader and cleaning rules of
``real_world_datasets.py:238-572`` and the agent partitioning of ``main.py:524-682`` (regular grid / k-d bisection / random /
sequential, optional per-agent subsampling), with the reference's names, argument meaning and error behaviour.

Also here, on the output side: ``evaluate_predictions`` (the metric dictionary of main.py:1598-1736, NLPD included), and the small
dataset helpers ``generate_data_numpy`` (main.py:457-522, the "classical dataset" mode) and ``save_quantum_dataset`` (main.py:433-455).

This is index and byte work on a few thousand rows; it stays in NumPy by design (the shards reach the GPU once, through the
engine's persistent pinned staging buffers).  Every function here selects exactly the rows the reference selects on the same
inputs - ``tests/test_data_plumbing.py`` compares with outputs of the real reference functions (``tests/golden/data_plumbing.npz``).
Printing and plotting of the reference are not reproduced.
"""
from __future__ import annotations

import os

import numpy as np

# real_world_datasets.py:265-289 (tile, bounds) and :447-452 (plausible elevation window, metres)
SRTM_REGIONS = {
    "maharashtra": ("N17E073", (17.0, 18.0, 73.0, 74.0), (0, 2000)),
    "great_lakes": ("N43W080", (43.0, 44.0, -80.0, -79.0), (75, 600)),
    "oregon_coast": ("N45W123", (45.0, 46.0, -123.0, -122.0), (0, 1500)),
    "washington_coast": ("N47W124", (47.0, 48.0, -124.0, -123.0), (0, 3000)),
}
HGT_SIDES = {3601 * 3601 * 2: 3601, 1201 * 1201 * 2: 1201}    # SRTM 1 and 3 arc-second tiles
HGT_NO_DATA = -32768


def get_tile_for_region(region):
    """real_world_datasets.py:574-582: tile name of a region, unknown names pass through."""
    return SRTM_REGIONS[region][0] if region in SRTM_REGIONS else region


def read_hgt_file(hgt_path):
    """real_world_datasets.py:527-572: a square tile of big-endian int16 heights, side chosen by the file size
    (3601 or 1201), returned as float64 with the no-data marker (-32768) left in place."""
    size = os.path.getsize(hgt_path)
    if size not in HGT_SIDES:
        raise ValueError(f"Unexpected HGT file size: {size} bytes")
    side = HGT_SIDES[size]
    return np.fromfile(hgt_path, dtype=">i2").reshape(side, side).astype(np.float64)


def load_srtm_elevation_dataset(region="maharashtra", max_samples=5000, subsample_factor=10, normalize=True, random_state=42,
                                save_plot=False, use_preprocessed=False, data_dir="srtm_data", preprocessed_dir="srtm/preprocessed"):
    """real_world_datasets.py:238-510: X = (lat, lon) of the kept grid points, Y = elevation.
    Rows are kept in row-major tile order (north to south, west to east) through: every ``subsample_factor``-th row/column;
    drop no-data / non-finite; drop negatives (all four regions are "no negative" regions, :266-289); keep the region's
    elevation window; if more than ``max_samples`` remain, ``np.random.seed(random_state)`` + ``choice`` without
    replacement; then coordinates min-max scaled to [-1, 1] and elevation standardised (population std).
    ``save_plot`` is accepted and ignored.  The directories are the reference's relative paths by default."""
    if region not in SRTM_REGIONS:
        raise ValueError(f"Region '{region}' not supported. Available: {list(SRTM_REGIONS.keys())}")
    tile, (lat_min, lat_max, lon_min, lon_max), (lo, hi) = SRTM_REGIONS[region]
    if use_preprocessed:
        path = os.path.join(preprocessed_dir, f"{tile}.npy")
        if not os.path.exists(path):
            raise FileNotFoundError(f"Preprocessed file not found: {path}")
        elevation = np.load(path)
        if elevation.ndim != 2 or elevation.shape[0] != elevation.shape[1]:
            raise ValueError(f"Unexpected preprocessed data shape: {elevation.shape}. Expected square grid.")
    else:
        candidates = [os.path.join(data_dir, f"{tile}.hgt"), os.path.join(data_dir, f"{tile}.SRTMGL1.hgt")]
        path = next((c for c in candidates if os.path.exists(c)), None)
        if path is None:
            raise FileNotFoundError(f"HGT file not found for tile {tile}\n  Looked for: {tile}.hgt or {tile}.SRTMGL1.hgt\n"
                                    f"  In directory: {os.path.abspath(data_dir)}")
        elevation = read_hgt_file(path)
    rows, cols = elevation.shape
    step = subsample_factor if subsample_factor > 1 else 1
    lat = np.linspace(lat_max, lat_min, rows)[::step]                  # north at the top
    lon = np.linspace(lon_min, lon_max, cols)[::step]
    y = elevation[::step, ::step].ravel()
    x = np.column_stack([np.repeat(lat, lon.size), np.tile(lon, lat.size)])
    keep = (y != HGT_NO_DATA) & np.isfinite(y)
    keep &= y >= 0
    keep &= (y >= lo) & (y <= hi)
    x, y = x[keep], y[keep]
    if len(y) > max_samples:
        np.random.seed(random_state)
        pick = np.random.choice(len(y), size=max_samples, replace=False)
        x, y = x[pick], y[pick]
    if normalize:
        x_min = x.min(axis=0, keepdims=True)
        x = 2.0 * (x - x_min) / (x.max(axis=0, keepdims=True) - x_min) - 1.0
        from sklearn.preprocessing import StandardScaler             # the reference's scaler (:492-493), same arithmetic
        y = StandardScaler().fit_transform(y.reshape(-1, 1)).flatten()
    return x, y


# ---------------------------------------------------------------------------------------------------------------------
def _kd_cells(points, n_cells):
    """main.py:524-553: split the largest cell (first one on ties) at the median of its widest coordinate (``<=`` goes left;
    the mean replaces the median if that leaves a side empty) until there are ``n_cells``; the left part takes the place of
    the split cell, the right part goes to the end of the list."""
    cells = [np.arange(len(points))]
    while len(cells) < n_cells:
        k = int(np.argmax([len(c) for c in cells]))
        cell = cells.pop(k)
        coord = points[cell][:, int(np.argmax(np.ptp(points[cell], axis=0)))]
        left = coord <= np.median(coord)
        if left.all() or not left.any():
            left = coord <= coord.mean()
        cells.insert(k, cell[left])
        cells.append(cell[~left])
    return cells


def _grid_cells(points, n_agents):
    """main.py:555-585: c^d equal boxes over the bounding box, agent id read as a base-c number with the FIRST coordinate
    most significant; both box faces are inclusive, so a point on an interior face belongs to both neighbours.  None when
    ``n_agents`` is not a perfect d-th power."""
    d = points.shape[1]
    c = round(n_agents ** (1 / d))
    if c ** d != n_agents:
        return None
    inside = []                                                        # inside[j][i]: rows whose coordinate j lies in slab i
    for j in range(d):
        edges = np.linspace(points[:, j].min(), points[:, j].max(), c + 1)
        inside.append([(points[:, j] >= edges[i]) & (points[:, j] <= edges[i + 1]) for i in range(c)])
    cells = []
    for agent in range(n_agents):
        digits = np.unravel_index(agent, (c,) * d)
        mask = np.ones(len(points), dtype=bool)
        for j in range(d):
            mask &= inside[j][digits[j]]
        cells.append(np.where(mask)[0])
    return cells


def sample_agent_data_percentage(X_agent, Y_agent, percentage, random_seed=42):
    """main.py:587-613: ``max(1, int(n * percentage))`` rows drawn without replacement after ``np.random.seed(random_seed)``."""
    if percentage <= 0.0 or percentage > 1.0:
        raise ValueError(f"Percentage must be between 0.0 and 1.0, got {percentage}")
    n = X_agent.shape[0]
    np.random.seed(random_seed)
    pick = np.random.choice(n, size=max(1, int(n * percentage)), replace=False)
    return X_agent[pick], Y_agent[pick]


def split_indices(X, n_agents, partition_method="regional", random_seed=42):
    """The row indices of every agent's shard (main.py:615-666) - the index form of ``split_data_numpy``."""
    n = X.shape[0]
    if partition_method == "regional":
        if (X.shape[1] if X.ndim > 1 else 1) == 1:
            return np.array_split(np.argsort(X[:, 0]), n_agents)
        cells = _grid_cells(X, n_agents)
        return cells if cells is not None else _kd_cells(X, n_agents)
    if partition_method == "random":
        np.random.seed(random_seed)
        return np.array_split(np.random.permutation(n), n_agents)
    if partition_method == "sequential":
        return np.array_split(np.arange(n), n_agents)
    raise ValueError(f"Unknown partition method: {partition_method}. Choose from: 'regional', 'random', 'sequential'")


def split_data_numpy(X, Y, n_agents, partition_method="regional", data_percentage=1.0, random_seed=42):
    """main.py:615-682: list of ``(X_agent, Y_agent)``; with ``data_percentage < 1`` every shard is subsampled with the SAME
    seed (``sample_agent_data_percentage``)."""
    shards = []
    for idx in split_indices(X, n_agents, partition_method, random_seed):
        xa, ya = X[idx], Y[idx]
        if data_percentage < 1.0:
            xa, ya = sample_agent_data_percentage(xa, ya, data_percentage, random_seed)
        shards.append((xa, ya))
    return shards


def prepare_training_data(X_full, Y_full, n_agents, partition_method="regional", test_split=0.1, seed=42, data_percentage=1.0,
                          split_seed=None):
    """main.py:2353-2372: scikit-learn ``train_test_split(shuffle=True, random_state=split_seed)`` (``split_seed`` is the
    data seed for the SRTM dataset, else ``seed`` - main.py:2355), then the agent partition of the training part with
    ``seed``.  Returns ``(shards, (X_train, Y_train), (X_test, Y_test))``; ``shards`` feeds ``run_admm`` / ``AdmmEngine``,
    the training pair is its ``cv_data``."""
    from sklearn.model_selection import train_test_split
    x_tr, x_te, y_tr, y_te = train_test_split(X_full, Y_full, test_size=test_split, shuffle=True,
                                              random_state=seed if split_seed is None else split_seed)
    return split_data_numpy(x_tr, y_tr, n_agents, partition_method, data_percentage, seed), (x_tr, y_tr), (x_te, y_te)


# ---------------------------------------------------------------------------------------------------------------------
def generate_data_numpy(num_samples, input_dim=1, noise_std=0.1, data_seed=None):
    """main.py:457-522, the ``--classical-dataset`` mode: inputs from ``np.random.seed(data_seed)`` + ``uniform``, a closed-form test
    function (1-D: the GRBCM paper's function on [0,1]; 2-D: normalised Goldstein-Price on [-2,2]^2; 3-D: negated Hartmann-3 on
    [0,1]^3) plus ``normal(0, noise_std)`` from the same stream.  ``data_seed=None`` = wall-clock seed, as in the reference."""
    if data_seed is None:
        import time
        data_seed = int(time.time() * 1000) % 2 ** 32
    np.random.seed(data_seed)
    if input_dim == 1:
        X = np.random.uniform(0, 1, size=(num_samples, 1))
        x = X[:, 0]
        Y = 5 * x ** 2 * np.sin(12 * x) + (x ** 3 - 0.5) * np.sin(3 * x - 0.5) + 4 * np.cos(2 * x)
    elif input_dim == 2:
        X = np.random.uniform(-2.0, 2.0, size=(num_samples, 2))
        u, v = X[:, 0], X[:, 1]
        first = 1 + (u + v + 1) ** 2 * (19 - 14 * u + 3 * u ** 2 - 14 * v + 6 * u * v + 3 * v ** 2)
        second = 30 + (2 * u - 3 * v) ** 2 * (18 - 32 * u + 12 * u ** 2 + 48 * v - 36 * u * v + 27 * v ** 2)
        Y = (np.log(first * second) - 8.693) / 2.427
    elif input_dim == 3:
        X = np.random.uniform(0.0, 1.0, size=(num_samples, 3))
        weights = np.array([1.0, 1.2, 3.0, 3.2])
        widths = np.array([[3.0, 10.0, 30.0], [0.1, 10.0, 35.0], [3.0, 10.0, 30.0], [0.1, 10.0, 35.0]])
        centres = 1e-4 * np.array([[3689.0, 1170.0, 2673.0], [4699.0, 4387.0, 7470.0], [1091.0, 8732.0, 5547.0], [381.0, 5743.0, 8828.0]])
        Y = np.zeros(num_samples)
        for w, a, c in zip(weights, widths, centres):
            Y += w * np.exp(-np.sum(a * (X - c) ** 2, axis=1))
        Y = -Y
    else:
        raise ValueError(f"Unsupported input dimension: {input_dim}")
    Y += np.random.normal(0, noise_std, num_samples)
    return X, Y


def save_quantum_dataset(X, Y, dataset_name, output_dir="quantum_datasets"):
    """main.py:433-455: one CSV ``<name>_<d>d_<n>.csv`` with the header ``X1,...,Xd,Y``; returns its path."""
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, f"{dataset_name}_{X.shape[1]}d_{X.shape[0]}.csv")
    header = ",".join([f"X{i + 1}" for i in range(X.shape[1])] + ["Y"])
    np.savetxt(path, np.column_stack((X, Y)), delimiter=",", header=header, comments="")
    return path


def _grade(value, cuts, labels):
    for cut, label in zip(cuts, labels):
        if value > cut:
            return label
    return labels[-1]


def evaluate_predictions(Y_true, Y_pred, Y_pred_var=None, dataset_type="Test"):
    """main.py:1598-1736 without the printing: mse, rmse, mae, r2, max_error, mean / std of the residuals, range-normalised rmse and
    the verbal grade; with variances also mean predictive std, 1-sigma / 2-sigma coverage, the uncertainty-normalised rmse, the mean NLPD
    (variance floored at 1e-10, main.py:1657-1662) and the calibration grade.  ``dataset_type`` only labelled the printout."""
    y, p = np.asarray(Y_true, dtype=np.float64), np.asarray(Y_pred, dtype=np.float64)
    res = y - p
    mse = float(np.mean(res ** 2))
    rmse = float(np.sqrt(mse))
    spread = float(np.sum((y - np.mean(y)) ** 2))
    r2 = 1.0 - float(np.sum(res ** 2)) / spread if spread > 0 else (1.0 if np.sum(res ** 2) == 0 else 0.0)
    y_range = y.max() - y.min()
    out = {"mse": mse, "rmse": rmse, "mae": float(np.mean(np.abs(res))), "r2": r2, "max_error": np.max(np.abs(res)),
           "mean_residual": np.mean(res), "std_residual": np.std(res),
           "normalized_rmse_range": rmse / y_range if y_range > 0 else float("inf"),
           "performance": _grade(r2, (0.9, 0.7, 0.5), ("Excellent", "Good", "Fair", "Poor"))}
    if Y_pred_var is not None:
        var = np.asarray(Y_pred_var, dtype=np.float64)
        std = np.sqrt(var)
        one, two = np.mean(np.abs(res) <= std), np.mean(np.abs(res) <= 2 * std)
        safe = np.maximum(var, 1e-10)
        nlpd = np.mean(0.5 * np.log(2 * np.pi) + 0.5 * np.log(safe) + 0.5 * (res ** 2 / safe))
        quality = "Good" if one > 0.5 and two > 0.8 else ("Fair" if one > 0.4 and two > 0.7 else "Poor")
        out.update({"mean_uncertainty": np.mean(std), "within_1sigma": one, "within_2sigma": two,
                    "normalized_rmse_uncertainty": np.sqrt(np.mean((res / std) ** 2)), "nlpd": nlpd, "uncertainty_quality": quality})
    return out
