"""Randomised soak of the persistent tensor-core training kernels (scripts/soak_train.py): random (B, T), both tiers, full and
half tiles, capped grids, eval and train mode -- every configuration must terminate (no mbarrier deadlock), be finite, and
give logits bit-identical to the uncapped launch."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_training_kernels_soak():
    r = subprocess.run([sys.executable, str(ROOT / "scripts" / "soak_train.py"), "40", "7"], cwd=ROOT, capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "soak ok" in r.stdout
