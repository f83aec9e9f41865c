"""ORACLE — CPU restatement of the reference's hot path (test infrastructure only; blocks and schedulers pinned against the
third-party package's published known-answer tests, tests/test_oracle_published_kats.py; the 0.18.2 inverse-scheduler pairing unpinned).

See oracle/unet.py, oracle/schedulers.py, oracle/pipeline.py for the per-function reference citations.
Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and --impl reference legs).
"""
from .pipeline import (OraclePipeline, oracle_custom_guided_generation, oracle_ddib, oracle_inversion,
                       oracle_linear_interp_custom_guidance_inverted_start)
from .schedulers import OracleDDIMInverseScheduler, OracleDDIMScheduler, rescale_zero_terminal_snr
from .unet import OracleCondUNet2D
