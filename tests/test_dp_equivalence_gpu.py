"""N-rank data-parallel step == 1-rank step on the concatenated batch, on real GPUs over NCCL (needs >= 2
devices; the 1-GPU test box skips it -- the recorded 2-GPU run is profiles/r02_dp_equivalence.jsonl)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('compute', ['fp32', 'bf16'])
def test_two_rank_step_equals_one_rank_step(compute, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    out = str(tmp_path / 'dp.json')
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tools', 'dp_equivalence.py'),
           '--compute', compute, '--out', out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(open(out).read().strip().splitlines()[-1])
    print(res)
    assert res['params_bit_identical_across_ranks']
    assert res['grad_rel_l2'] < (1e-5 if compute == 'fp32' else 2e-2)
    assert abs(res['loss_dp_mean'] - res['loss_1rank']) < 1e-5 * max(1.0, abs(res['loss_1rank'])) * (1 if compute == 'fp32' else 100)
