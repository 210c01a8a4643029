"""dqgp_b200 — B200-native engine for the hot path of distributed quantum Gaussian-process regression.

Host-side mirror of the reference's interface for that path (same names, argument meaning and error
behaviour as ``main.py`` / ``agent_riemannian.py`` / ``riemannian_optimizer.py`` of
mpala-lab/distributed-quantum-gaussian-processes) over the C ABI of ``include/dqgp.h`` (hand-written
sm_100a CUDA in ``csrc/``).  PyTorch is used for device memory, streams and ``torch.distributed`` only.
Import as ``dqgp_b200`` (the repo-root shim) — the directory name carries the reference's repository name.
"""
from ._lib import DqgpError, load  # noqa: F401
from .riemannian import (RiemannianADMM, RiemannianOptimizer, TorusManifold, circular_mean,  # noqa: F401
                         create_riemannian_framework)
from .kernels import (EncodingCircuit, Executor, FidelityKernel, ProjectedQuantumKernel,  # noqa: F401
                      create_quantum_kernel)
from .agent import RiemannianAgent, process_agent_training, train_agents  # noqa: F401
from .engine import AgentEngine, AdmmEngine, agent_block, exchange_rows, synthetic_dataset  # noqa: F401
from .datagen import generate_quantum_gp_data  # noqa: F401
from .data import (evaluate_predictions, generate_data_numpy, get_tile_for_region, load_srtm_elevation_dataset,  # noqa: F401
                   prepare_training_data, read_hgt_file, sample_agent_data_percentage, save_quantum_dataset, split_data_numpy,
                   split_indices)
from .driver import run_admm  # noqa: F401
from .predict import k_fold_cross_validation_consensus, nlpd, predict_quantum_gp  # noqa: F401

__version__ = "0.1.0"
