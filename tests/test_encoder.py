"""The reference's full backbone: frozen MiniLM-L6 BertModel forward on the device
against the oracle SURVEY.md section 8c names — transformers' own `BertModel(BertConfig(30522, 384, 6, 12, 1536))`
with the same (random-init) weights, fp32 on the CPU.  transformers is a third-party dependency of the reference
(backend/uv.lock pins 4.52.4; 5.5 is installed), not part of /root/reference."""
import pytest
import torch


def _hf_twin(backbone):
    from transformers import BertConfig, BertModel

    cfg = backbone.config
    hf = BertModel(BertConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                              num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                              intermediate_size=cfg.intermediate_size,
                              max_position_embeddings=cfg.max_position_embeddings)).eval()
    missing, unexpected = hf.load_state_dict({k: v.detach().cpu() for k, v in backbone.state_dict().items()}, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in k or "token_type_ids" in k for k in missing), missing
    return hf


def test_backbone_has_the_reference_parameter_counts_and_hf_key_names():
    """models/e15.lr4.d384.m3_summary.txt:7-9: BertEmbeddings 11,918,592 / BertEncoder 10,646,784 / BertPooler 147,840;
    every parameter name is one transformers' BertModel also has (a real checkpoint loads)."""
    from two_towers_overlords_b200.encoder import MiniLMBackbone

    m = MiniLMBackbone()
    count = lambda mod: sum(p.numel() for p in mod.parameters())  # noqa: E731
    assert (count(m.embeddings), count(m.encoder), count(m.pooler)) == (11_918_592, 10_646_784, 147_840)
    assert not any(p.requires_grad for p in m.parameters())
    small = MiniLMBackbone(vocab_size=100, hidden_size=128, num_hidden_layers=2, num_attention_heads=4,
                           intermediate_size=256, max_position_embeddings=32)
    _hf_twin(small)  # raises on a name / shape mismatch


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(6, 32), (5, 47), (3, 256)])
def test_minilm_forward_matches_transformers_bert(shape):
    """Last hidden state of every unmasked token and the pooled, normalised tower input: 1e-4 relative (north_star's
    gate for pooled embeddings) against BertModel on the CPU, ragged lengths, one fully padded tail."""
    import two_towers_overlords_b200 as tt
    from two_towers_overlords_b200.encoder import MiniLMBackbone

    torch.manual_seed(3)
    B, L = shape
    m = MiniLMBackbone().cuda()
    hf = _hf_twin(m)
    g = torch.Generator().manual_seed(B * 1000 + L)
    ids = torch.randint(999, 30522, (B, L), generator=g)
    lens = torch.randint(3, L + 1, (B,), generator=g)
    lens[0] = L
    mask = (torch.arange(L)[None, :] < lens[:, None]).long()
    ids = ids * mask
    with torch.no_grad():
        want = hf(input_ids=ids, attention_mask=mask)[0]
    got = m(input_ids=ids.cuda(), attention_mask=mask.cuda())[0].cpu()
    valid = mask.bool()
    err = ((got - want)[valid].norm(dim=-1) / want[valid].norm(dim=-1)).max()
    assert float(err) < 1e-4, float(err)
    # the tower's pooled input through the same masked-mean + normalise kernel (model.py:55-56)
    tower = tt.model.AveragePoolingTower(projection_dim=64, backbone="minilm").cuda()
    tower.pretrained_model.load_state_dict(m.state_dict())
    pooled = tower.pooled((ids.cuda(), mask.cuda())).cpu()
    mm = mask.unsqueeze(-1).float()
    ref = torch.nn.functional.normalize((want * mm).sum(1) / mm.sum(1).clamp(min=1e-9), p=2, dim=1)
    assert float(((pooled - ref).norm(dim=1) / ref.norm(dim=1)).max()) < 1e-4


@pytest.mark.gpu
def test_tower_with_minilm_backbone_trains_its_projection():
    """The reference-shaped step (model.py:40-60 + TripletLoss + backward) on top of the frozen encoder: gradients
    reach the projection only, and equal autograd's through the same pooled inputs."""
    import two_towers_overlords_b200 as tt

    torch.manual_seed(0)
    model = tt.TwoTowersModel(projection_dim=64, backbone="minilm").cuda()
    g = torch.Generator().manual_seed(1)
    mk = lambda B, L: (torch.randint(999, 30522, (B, L), generator=g).cuda(), torch.ones(B, L, dtype=torch.long).cuda())  # noqa: E731
    q, p, n = mk(8, 12), mk(8, 40), mk(8, 40)
    loss = tt.TripletLoss(0.3)(model.encode_queries(q), model.encode_documents(p), model.encode_documents(n))
    loss.backward()
    trainable = [nm for nm, prm in model.named_parameters() if prm.grad is not None]
    assert sorted(trainable) == sorted(f"{t}.projection.{i}.{k}" for t in ("query_tower", "document_tower")
                                       for i in (0, 2) for k in ("weight", "bias"))
    # same loss / gradients from plain torch on the pooled inputs
    xq, xp, xn = (model.query_tower.pooled(q), model.document_tower.pooled(p), model.document_tower.pooled(n))
    ref_params = [prm.detach().clone().requires_grad_(True) for prm in model.projection_parameters()]

    def mlp(x, w):
        return torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(x, w[0], w[1])), w[2], w[3])

    yq, yp, yn = mlp(xq, ref_params[:4]), mlp(xp, ref_params[4:]), mlp(xn, ref_params[4:])
    cd = lambda a, b: 1 - torch.nn.functional.cosine_similarity(a, b, dim=1)  # noqa: E731
    ref_loss = torch.relu(cd(yq, yp) - cd(yq, yn) + 0.3).mean()
    ref_loss.backward()
    assert abs(float(loss.detach()) - float(ref_loss.detach())) < 1e-4 * abs(float(ref_loss.detach()))
    for prm, ref in zip(model.projection_parameters(), ref_params):
        assert float((prm.grad - ref.grad).norm() / ref.grad.norm()) < 2e-3
