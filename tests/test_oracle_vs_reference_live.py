"""Live pin of the oracle against the reference's OWN modules on fresh seeds (beyond the committed golden vectors).
Runs only where /root/reference exists (the build container); skipped on the GPU box.  The reference is imported in a
subprocess because it patches transformers.AutoModel / AutoTokenizer process-wide (oracle/gen_golden.py:_install)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/backend"

SCRIPT = textwrap.dedent(
    """
    import sys, numpy as np, torch
    sys.path.insert(0, %(root)r)
    from oracle import gen_golden as G
    from oracle import two_towers_oracle as O
    G._install()
    import model as ref_model          # the reference's backend/model.py, unmodified
    import training as ref_training    # the reference's backend/training.py, unmodified

    worst = 0.0
    for seed, P, B, margin in ((1, 24, 9, 0.3), (2, 48, 17, 0.2), (3, 128, 5, 0.5)):
        rng = np.random.default_rng(seed)
        G.BANK.clear()
        G._make_bank(rng, 40, 90)
        torch.manual_seed(seed)
        m = ref_model.TwoTowersModel(projection_dim=P)
        o = O.OracleTwoTowers(P, vocab=G.V, hidden=G.H)
        with torch.no_grad():
            for tr, to in ((m.query_tower, o.query_tower), (m.document_tower, o.document_tower)):
                to.table.copy_(tr.pretrained_model.emb.weight)
                for i in (0, 2):
                    to.projection[i].weight.copy_(tr.projection[i].weight)
                    to.projection[i].bias.copy_(tr.projection[i].bias)
        crit = ref_model.TripletLoss(margin=margin)
        opt_r = torch.optim.Adam(m.parameters(), lr=1e-3)
        opt_o = torch.optim.Adam(o.parameters(), lr=1e-3)
        for step in range(3):
            qs = [f"q:{i}" for i in rng.integers(0, 40, B)]
            ps = [f"d:{i}" for i in rng.integers(0, 90, B)]
            ns = [f"d:{i}" for i in rng.integers(0, 90, B)]
            avg = ref_training.train_epoch(m, [(qs, ps, ns)], crit, opt_r, log_wandb=False)
            toks = []
            for texts in (qs, ps, ns):
                ids, msk = G._tok(texts)
                toks += [torch.from_numpy(ids), torch.from_numpy(msk)]
            loss = O.train_step(o, opt_o, tuple(toks), margin)
            assert abs(loss - avg) <= 1e-6 * max(1.0, abs(avg)), (seed, step, loss, avg)
            for tr, to in ((m.query_tower, o.query_tower), (m.document_tower, o.document_tower)):
                for i in (0, 2):
                    d = (tr.projection[i].weight - to.projection[i].weight).abs().max().item()
                    worst = max(worst, d)
                    assert d <= 1e-6, (seed, step, i, d)
        # forward of the trained models still agrees
        q_ref = m.encode_queries(qs).detach()
        ids, msk = G._tok(qs)
        q_or = o.encode_queries(torch.from_numpy(ids), torch.from_numpy(msk)).detach()
        assert (q_ref - q_or).abs().max().item() <= 1e-5
    print("LIVE_PIN_OK", worst)
    """
)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_oracle_matches_the_reference_modules_on_fresh_seeds():
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=900,
                         env={**os.environ, "HF_HUB_OFFLINE": "1", "WANDB_MODE": "disabled"})
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert "LIVE_PIN_OK" in out.stdout
