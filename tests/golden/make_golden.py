#!/usr/bin/env python
"""Generate tests/golden/*.pt by running the REFERENCE's own modules on the CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so the oracle
(oracle/) and the CUDA path are pinned against outputs of the reference's own
Python code, produced here with fixed seeds and committed as small fixtures.
Nothing under tests/, bench.py or smoke() reads /root/reference at run time.

Imports follow SURVEY.md section 8c: tower_code/* import as-is with their
directory on sys.path; item_tower.py needs stubs for `sqlalchemy`, `database`
and an offline BertConfig() model (here a SMALL BertConfig so that the fixture
stays a few MB -- the code path, `bert_model.embeddings(input_ids=...)`, is the
same HF module).  Inline (function-less) reference code -- the SimCSE loss
item_tower.py:1075-1082 and the retrieval lines v1_usertower_train.py:672-675,
mined_inference.py:1536-1542 -- is executed by slicing those very source lines
out of the reference files and exec()-ing them, so that what is pinned is the
reference's text, not a paraphrase.
"""
import os
import sys
import textwrap
import types
from types import SimpleNamespace

import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def ref_lines(relpath, lo, hi):
    with open(os.path.join(REF, relpath), encoding="utf-8") as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[lo - 1:hi]))


def save(name, obj):
    path = os.path.join(OUT, name)
    torch.save(obj, path)
    print(f"wrote {name}: {os.path.getsize(path) / 1e6:.2f} MB")


def import_tower_code():
    sys.path.insert(0, os.path.join(REF, "tower_code"))
    import v1_refine_usertower as ru
    import v1_usertower_train as ut
    import mined_inference as mi
    return ru, ut, mi


def import_item_tower():
    """3-part shim of SURVEY.md 8c."""
    sa = types.ModuleType("sqlalchemy")
    sa.select = lambda *a, **k: None
    sys.modules["sqlalchemy"] = sa
    from pydantic import BaseModel
    from typing import Dict, Any
    db = types.ModuleType("database")

    class TrainingItem(BaseModel):
        product_id: str
        feature_data: Dict[str, Any]
        product_name: str

    class ProductInferenceInput:  # noqa: D401 - dummy ORM class
        pass

    db.TrainingItem, db.ProductInferenceInput = TrainingItem, ProductInferenceInput
    sys.modules["database"] = db
    import transformers
    from transformers import BertConfig, BertModel
    small = dict(vocab_size=600, hidden_size=768, num_hidden_layers=1, num_attention_heads=12,
                 intermediate_size=256, max_position_embeddings=40)
    transformers.AutoConfig.from_pretrained = staticmethod(lambda *a, **k: BertConfig(**small))
    transformers.AutoModel.from_pretrained = staticmethod(lambda *a, **k: BertModel(BertConfig(**small)))
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.makedirs("/tmp/_golden_cwd", exist_ok=True)
    os.chdir("/tmp/_golden_cwd")          # item_tower.py:31 does os.makedirs("models")
    try:
        import item_tower as it
    finally:
        os.chdir(cwd)
    return it


# --------------------------------------------------------------------------
def make_user_tower(ru, ut):
    torch.manual_seed(42)
    args = SimpleNamespace(d_model=128, max_len=10, dropout=0.2, pretrained_dim=128, nhead=4, num_layers=2,
                           num_items=300, num_prod_types=40, num_colors=30, num_graphics=20, num_sections=25)
    model = ru.SASRecUserTower(args)
    with torch.no_grad():                       # make gates distinct so that a gate mix-up is caught
        model.seq_gate.copy_(torch.linspace(-0.5, 0.7, 6))
        model.static_gate.copy_(torch.linspace(0.9, -0.6, 10))
    model.eval()                                # dropout off: deterministic
    B, L = 6, 10
    g = torch.Generator().manual_seed(7)
    lens = torch.tensor([10, 7, 1, 4, 10, 2])
    pad = torch.arange(L).unsqueeze(0) < (L - lens).unsqueeze(1)          # left padding
    def ids(hi):
        x = torch.randint(1, hi + 1, (B, L), generator=g)
        return x.masked_fill(pad, 0)
    inp = dict(
        pretrained_vecs=F.normalize(torch.randn(B, L, 128, generator=g), dim=-1) * (~pad).unsqueeze(-1),
        item_ids=ids(300), time_bucket_ids=ids(9), type_ids=ids(40), color_ids=ids(30),
        graphic_ids=ids(20), section_ids=ids(25),
        age_bucket=torch.randint(0, 11, (B,), generator=g), price_bucket=torch.randint(0, 11, (B,), generator=g),
        cnt_bucket=torch.randint(0, 11, (B,), generator=g), recency_bucket=torch.randint(0, 11, (B,), generator=g),
        channel_ids=torch.randint(0, 4, (B,), generator=g), club_status_ids=torch.randint(0, 4, (B,), generator=g),
        news_freq_ids=torch.randint(0, 3, (B,), generator=g), fn_ids=torch.randint(0, 3, (B,), generator=g),
        active_ids=torch.randint(0, 3, (B,), generator=g), cont_feats=torch.randn(B, 4, generator=g),
        padding_mask=pad)
    cap = {}
    h1 = model.emb_ln.register_forward_hook(lambda m, i, o: cap.__setitem__("seq_front", i[0].detach().clone()))
    h2 = model.static_mlp.register_forward_hook(lambda m, i, o: cap.__setitem__("static_front", i[0].detach().clone()))
    out_train = model(**inp, training_mode=True)
    h1.remove(); h2.remove()
    w = torch.randn(out_train.shape, generator=g)
    (out_train * w).sum().backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    # NB no torch.no_grad() here: under no_grad + eval() nn.TransformerEncoder takes ATen's fused "fast path",
    # which on CPU returns NaN for every left-padded user (fully masked query rows); with grad enabled the
    # ordinary path is taken.  That is stock-ATen behaviour outside the hot path, so the fixture avoids it.
    out_eval = model(**inp, training_mode=False).detach()
    save("user_tower.pt", dict(args=vars(args), state=model.state_dict(), inputs=inp, out_train=out_train.detach(),
                               out_eval=out_eval, seq_front=cap["seq_front"], static_front=cap["static_front"],
                               cotangent=w, grads=grads))

    # SASRecItemTower (U4): normalize-whole-table-then-gather, as the loop does (:810-811 + loss :833)
    torch.manual_seed(3)
    logq = torch.log(torch.rand(301) + 1e-6); logq[0] = -20.0
    it = ut.SASRecItemTower(300, 128, log_q_tensor=logq)
    pre = F.normalize(torch.randn(301, 128), dim=1); pre[0] = 0
    it.init_from_pretrained(pre)
    tgt = torch.randint(0, 301, (40,), generator=g)
    norm_all = F.normalize(it.get_all_embeddings(), p=2, dim=1)
    rows = norm_all[tgt]
    cot = torch.randn(40, 128, generator=g)
    (rows * cot).sum().backward()
    save("item_matrix.pt", dict(state=it.state_dict(), target_ids=tgt, rows=rows.detach(), cotangent=cot,
                                grad_weight=it.item_matrix.weight.grad.clone(), log_q=it.get_log_q().clone()))


def make_losses(ru, mi, it_mod):
    g = torch.Generator().manual_seed(11)
    N, D, V = 96, 128, 50
    U = F.normalize(torch.randn(N, D, generator=g), dim=1)
    table = F.normalize(torch.randn(V, D, generator=g), dim=1)
    # correlate users with their targets so that logits are not all alike
    tgt = torch.randint(0, V, (N,), generator=g)            # collisions guaranteed (N > V), some zeros
    U = F.normalize(U + 1.5 * table[tgt], dim=1)
    uid = torch.randint(0, 30, (N,), generator=g)
    logq = torch.log(torch.rand(V, generator=g) + 1e-6); logq[0] = -20.0
    probs = torch.rand(V, generator=g)
    U2 = F.normalize(U + 0.3 * torch.randn(N, D, generator=g), dim=1)
    out = dict(U=U, U2=U2, table=table, tgt=tgt, uid=uid, logq=logq, probs=probs)

    def run(name, fn, wrt, *a, **k):
        leaves = [x.clone().requires_grad_(True) for x in wrt]
        r = fn(*leaves, *a, **k)
        stats = None
        if isinstance(r, tuple):
            r, stats = r
        r.backward()
        out[name] = dict(loss=r.detach(), grads=[x.grad.clone() for x in leaves], stats=stats)

    run("c2", ru.inbatch_corrected_logq_loss, [U, table], tgt, uid, logq, temperature=0.1, lambda_logq=1.0)
    run("c2_nologq", ru.inbatch_corrected_logq_loss, [U, table], tgt, uid, logq, temperature=0.07, lambda_logq=0.0)
    run("c3", ru.duorec_loss_refined, [U, U2], tgt, temperature=0.1, lambda_sup=0.1)
    run("c3_nosup", ru.duorec_loss_refined, [U, U2], tgt, temperature=0.1, lambda_sup=0.0)
    tgt_nopos = torch.arange(1, N + 1) % V; tgt_nopos = torch.arange(N) + 1   # all distinct -> mask.sum()==0 branch
    out["tgt_distinct"] = tgt_nopos
    run("c3_distinct", ru.duorec_loss_refined, [U, U2], tgt_nopos, temperature=0.1, lambda_sup=0.1)
    run("c4", ru.full_batch_hard_emphasis_loss, [U, table], tgt, logq, top_k_percent=0.05, hard_margin=0.01,
        hnm_threshold=0.90, temperature=0.15, lambda_logq=1.0)
    run("c5_hnm", ru.inbatch_hnm_corrected_loss_with_stats, [U, table], tgt, logq, top_k_percent=0.05,
        hnm_threshold=0.90, temperature=0.1, lambda_logq=0.7)
    torch.manual_seed(1234)                    # the function draws torch.randint(0, N, (N, 100)) internally (:722)
    out["mixed_random_indices"] = torch.randint(0, N, (N, 100))
    torch.manual_seed(1234)
    run("c5_mixed", ru.inbatch_mixed_hnm_loss_with_stats, [U, table], tgt, logq, top_k_percent=0.05)
    Vrows = table[tgt]
    run("c5_logq", mi.logq_correction_loss, [U, Vrows], tgt, probs, temperature=0.07, lambda_logq=0.5)
    run("c5_eff", mi.efficient_corrected_logq_loss, [U, Vrows], tgt, logq, temperature=0.1, lambda_logq=0.1)

    # C1: the inline SimCSE loss, item_tower.py:1075-1082, executed from the reference's source text
    src = ref_lines("item_tower.py", 1075, 1082)
    E1 = F.normalize(torch.randn(64, D, generator=g), dim=1)
    E2 = F.normalize(E1 + 0.4 * torch.randn(64, D, generator=g), dim=1)
    e1, e2 = E1.clone().requires_grad_(True), E2.clone().requires_grad_(True)
    ns = dict(torch=torch, emb1=e1, emb2=e2, DEVICE="cpu", loss_func=torch.nn.CrossEntropyLoss())
    exec(src, ns)
    ns["loss"].backward()
    out["c1"] = dict(E1=E1, E2=E2, loss=ns["loss"].detach(), temperature=ns["temperature"],
                     grads=[e1.grad.clone(), e2.grad.clone()], source=src)
    save("losses.pt", out)


def make_retrieval():
    g = torch.Generator().manual_seed(5)
    nu, ni, D = 48, 2500, 128
    Uq = F.normalize(torch.randn(nu, D, generator=g), dim=1)
    I = F.normalize(torch.randn(ni, D, generator=g), dim=1)
    out = dict(U=Uq, I=I)
    # v1_usertower_train.py:672-675
    src = ref_lines("tower_code/v1_usertower_train.py", 672, 675)
    for k in (12, 20, 100, 500):
        ns = dict(torch=torch, valid_user_emb=Uq, norm_item_embeddings=I, max_k=k)
        exec(src, ns)
        sc, idx = torch.topk(ns["scores"], k=k, dim=-1)
        assert torch.equal(idx, ns["topk_indices"])
        out[f"k{k}"] = dict(ids=ns["topk_indices"], scores=sc)
    out["source"] = src
    # mined_inference.py:1536-1542 (GNN variant masks item 0)
    src2 = ref_lines("tower_code/mined_inference.py", 1536, 1542)
    ns = dict(torch=torch, batch_gnn_user_norm=Uq, all_gnn_items_norm=I, max_k=20)
    exec(src2, ns)
    sc, _ = torch.topk(ns["scores"], k=20, dim=1)
    out["gnn_k20"] = dict(ids=ns["topk_indices"], scores=sc)
    out["source_gnn"] = src2
    # ties: duplicate item rows -> equal scores; the fixture records the score multiset
    I_t = I.clone(); I_t[100] = I_t[7]; I_t[2000] = I_t[7]
    scores = Uq @ I_t.T
    sc, idx = torch.topk(scores, k=12, dim=-1)
    out["ties_k12"] = dict(I=I_t, ids=idx, scores=sc)
    save("retrieval.pt", out)


def make_item_front(it):
    torch.manual_seed(0)
    from utils import vocab
    V = vocab.get_std_vocab_size()
    enc = it.HybridItemTower(std_vocab_size=V, num_std_fields=6)
    enc.eval()
    enc._debug_logged = True
    B = 3
    g = torch.Generator().manual_seed(9)
    std = torch.randint(0, V, (B, 6), generator=g); std[0, 2] = 0; std[2, 0] = 1
    T = 32
    re_ids = torch.zeros(B, 9, T, dtype=torch.long); re_mask = torch.zeros(B, 9, T, dtype=torch.long)
    for b in range(B):
        for f in range(9):
            n = 0 if torch.rand((), generator=g) < 0.3 else int(torch.randint(3, 12, (), generator=g))
            if n:
                re_ids[b, f, :n] = torch.randint(3, 600, (n,), generator=g); re_mask[b, f, :n] = 1
    txt_ids = torch.zeros(B, T, dtype=torch.long); txt_mask = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(4, 16, (), generator=g))
        txt_ids[b, :n] = torch.randint(3, 600, (n,), generator=g); txt_mask[b, :n] = 1
    cap = {}
    hs = [enc.std_ln.register_forward_hook(lambda m, i, o: cap.update(std_pre=i[0].detach().clone(), std_out=o.detach().clone())),
          enc.re_ln.register_forward_hook(lambda m, i, o: cap.update(re_out=o.detach().clone())),
          enc.re_proj.register_forward_hook(lambda m, i, o: cap.update(word_embs=i[0].detach().clone())),
          ]
    out = enc(std, re_ids, re_mask, txt_ids, txt_mask)
    for h in hs:
        h.remove()
    w = torch.randn(out.shape, generator=g)
    (out * w).sum().backward()
    emb = enc.bert_model.embeddings
    keep = {k: v.detach().clone() for k, v in enc.state_dict().items()
            if k.startswith(("std_", "re_", "bert_model.embeddings."))}
    save("item_front.pt", dict(
        state=keep, inputs=dict(std_input=std, re_input_ids=re_ids, re_attn_mask=re_mask,
                                text_input_ids=txt_ids, text_attn_mask=txt_mask),
        std_out=cap["std_out"], word_embs=cap["word_embs"], re_out=cap["re_out"], out=out.detach(), cotangent=w,
        grad_std_embedding=enc.std_embedding.weight.grad.clone(),
        grad_std_field_emb=enc.std_field_emb.grad.clone(),
        grad_word_embeddings=emb.word_embeddings.weight.grad.clone(),
        bert_ln_eps=float(emb.LayerNorm.eps), vocab_size=V))


def make_hybrid_user(mi):
    torch.manual_seed(1)
    nu, ni = 400, 200
    gu = torch.randn(nu, 64); gi = torch.randn(ni, 64); gi[0] = 0
    ic = F.normalize(torch.randn(ni, 128), dim=1); ic[0] = 0
    m = mi.HybridUserTower(nu, ni, gu, gi, ic)
    m.eval()
    g = torch.Generator().manual_seed(2)
    B, L = 7, 9
    lens = torch.randint(1, L + 1, (B,), generator=g)
    seq_mask = (torch.arange(L).unsqueeze(0) < lens.unsqueeze(1)).long()       # right padded (pad_sequence)
    seq_ids = torch.randint(1, ni, (B, L), generator=g) * seq_mask
    deltas = torch.randint(0, 3000, (B, L), generator=g) * seq_mask             # exercises clamp(max=1000)
    u_idx = torch.randint(0, nu, (B,), generator=g)
    u_dense = torch.randn(B, 3, generator=g); u_cat = torch.randint(0, 2, (B,), generator=g)
    cap = {}
    hs = []
    for name in ("gnn_user_emb", "gnn_item_emb", "item_content_emb", "time_emb", "channel_emb"):
        hs.append(getattr(m, name).register_forward_hook(
            lambda mod, i, o, name=name: cap.__setitem__(name, o.detach().clone())))
    out, v_seq, gate = m(u_idx, seq_ids, deltas, seq_mask, u_dense, u_cat)
    for h in hs:
        h.remove()
    w = torch.randn(out.shape, generator=g)
    (out * w).sum().backward()
    tabs = {n: getattr(m, n).weight.detach().clone() for n in cap}
    grads = {n: (getattr(m, n).weight.grad.clone() if getattr(m, n).weight.grad is not None else None) for n in cap}
    save("hybrid_user.pt", dict(tables=tabs, inputs=dict(u_idx=u_idx, seq_ids=seq_ids, seq_deltas=deltas,
                                                          seq_mask=seq_mask, u_dense=u_dense, u_cat=u_cat),
                                gathered=cap, grads=grads, out=out.detach()))


def make_alignment(ut):
    """N2: the on-disk contract (utils/inference_utils.py:84-85,200-202: `pretrained_item_matrix.pt` = [N, 128] fp32,
    `item_ids.pt` = list[str]) read back by load_aligned_pretrained_embeddings (v1_usertower_train.py:131-160)."""
    import tempfile
    g = torch.Generator().manual_seed(17)
    n_pre, n_cur, D = 700, 900, 128
    pre = F.normalize(torch.randn(n_pre, D, generator=g), dim=1)
    # exported ids: zero-padded numeric strings (H&M article ids), sorted by str as the exporter does; one id twice
    # (the reference's dict keeps the LAST occurrence)
    pool = [f"{int(x):010d}" for x in torch.randperm(5000, generator=g)[:1200] + 108775000]
    pre_ids = sorted(pool[:n_pre - 1]) + [pool[3]]
    # the processor's catalogue: overlaps the export partly, in its own order
    perm = torch.randperm(1200, generator=g).tolist()
    cur_ids = [pool[i] for i in perm[:n_cur]]
    out = dict(pretrained=pre, pretrained_ids=pre_ids, item_ids=cur_ids, dim=D, seed=123)
    with tempfile.TemporaryDirectory() as d:
        torch.save(pre, os.path.join(d, "pretrained_item_matrix.pt"))
        torch.save(pre_ids, os.path.join(d, "item_ids.pt"))
        proc = SimpleNamespace(num_items=n_cur, item_ids=cur_ids)
        torch.manual_seed(123)
        out["aligned"] = ut.load_aligned_pretrained_embeddings(proc, d, D)
        # dict-wrapped tensor + tensor ids (both accepted by the reference, :141-146)
        torch.save({"weight": pre}, os.path.join(d, "pretrained_item_matrix.pt"))
        int_ids = torch.tensor([int(x) for x in pre_ids])
        torch.save(int_ids, os.path.join(d, "item_ids.pt"))
        proc2 = SimpleNamespace(num_items=n_cur, item_ids=[str(int(x)) for x in cur_ids])
        torch.manual_seed(123)
        out["aligned_int_ids"] = ut.load_aligned_pretrained_embeddings(proc2, d, D)
        out["pretrained_int_ids"] = int_ids
        # missing files -> random init only (:157-158)
        torch.manual_seed(123)
        out["aligned_missing"] = ut.load_aligned_pretrained_embeddings(proc, os.path.join(d, "nope"), D)
    save("alignment.pt", out)


def make_ensemble():
    """N4: the candidate-union ensembles of mined_inference.py -- min-max weighted sum (:1103-1183) and weighted RRF
    (:1347-1407) -- by exec()-ing the reference's own lines on synthetic normalised vectors."""
    import numpy as np
    g = torch.Generator().manual_seed(23)
    b, n, d1, d2, pool_k, max_k = 24, 1500, 64, 128, 100, 50
    ug = F.normalize(torch.randn(b, d1, generator=g), dim=1)
    ig = F.normalize(torch.randn(n, d1, generator=g), dim=1)
    us = F.normalize(torch.randn(b, d2, generator=g), dim=1)
    it_ = F.normalize(torch.randn(n, d2, generator=g) + 0.0, dim=1)
    alphas = [0.0, 0.3, 0.5, 0.7, 1.0]
    out = dict(user_gnn=ug, items_gnn=ig, user_seq=us, items_seq=it_, pool_k=pool_k, max_k=max_k, alphas=alphas, k_rrf=60)
    ns = dict(torch=torch, F=F, np=np, user_gnn_vecs=ug, all_gnn_item_vecs=ig, user_seq_vecs=us, all_seq_item_vecs=it_,
              pool_k=pool_k, current_batch_size=b, max_k=max_k, device="cpu", k_rrf=60)
    exec(ref_lines("tower_code/mined_inference.py", 1103, 1108), ns)        # two global top-M
    exec(ref_lines("tower_code/mined_inference.py", 1115, 1115), ns)        # union
    exec(ref_lines("tower_code/mined_inference.py", 1124, 1133), ns)        # gather + re-score
    exec(ref_lines("tower_code/mined_inference.py", 1139, 1145), ns)        # min-max
    out["combined_indices"] = ns["combined_indices"]
    out["s_gnn"], out["s_seq"] = ns["s_gnn"], ns["s_seq"]
    cic = ns["combined_indices"].cpu().numpy()

    def dedup(local_topk):
        rows = []
        for i in range(b):
            env = dict(np=np, pred_global_ids=cic[i][local_topk[i]])
            exec(ref_lines("tower_code/mined_inference.py", 1182, 1183), env)
            rows.append(torch.from_numpy(env["pred_unique"].copy()))
        return rows
    mm = {}
    for alpha in alphas:
        ns["alpha"] = alpha
        exec(ref_lines("tower_code/mined_inference.py", 1162, 1162), ns)    # weighted sum
        exec(ref_lines("tower_code/mined_inference.py", 1170, 1171), ns)    # topk(max_k + 20) -> numpy
        mm[alpha] = dict(final_scores=ns["final_scores"].clone(), pred_unique=dedup(ns["local_topk_indices"]))
    out["minmax"] = mm
    # RRF
    exec(ref_lines("tower_code/mined_inference.py", 1347, 1348), ns)        # raw scores
    exec(ref_lines("tower_code/mined_inference.py", 1358, 1371), ns)        # ranks by double sort + scatter
    exec(ref_lines("tower_code/mined_inference.py", 1379, 1380), ns)        # rrf scores
    out["rank_gnn"], out["rank_seq"] = ns["rank_gnn"], ns["rank_seq"]
    rr = {}
    for alpha in alphas:
        ns["alpha"] = alpha
        exec(ref_lines("tower_code/mined_inference.py", 1393, 1393), ns)
        exec(ref_lines("tower_code/mined_inference.py", 1396, 1397), ns)
        rr[alpha] = dict(final_scores=ns["final_rrf_scores"].clone(), pred_unique=dedup(ns["local_topk_indices"]))
    out["rrf"] = rr
    save("ensemble.pt", out)


def make_lightgcl():
    """N3: LightGCL's BPR / InfoNCE / L2 terms (gnn_model/v1_lightgcl.py:188-222) and its retrieval lines
    (gnn_model/v1_evaluate_lightgcl.py:312-318), called on the reference class with a stand-in `self`."""
    sys.path.insert(0, os.path.join(REF, "gnn_model"))
    import v1_lightgcl as lg
    g = torch.Generator().manual_seed(31)
    n_users, n_items, D, B = 300, 500, 64, 256
    n = n_users + n_items
    local = (torch.randn(n, D, generator=g) * 0.3).requires_grad_(True)
    glob = (torch.randn(n, D, generator=g) * 0.3).requires_grad_(True)
    users = torch.randint(0, n_users, (B,), generator=g)
    pos = torch.randint(n_users, n, (B,), generator=g)
    neg = torch.randint(n_users, n, (B,), generator=g)
    me = SimpleNamespace(temp=0.2)
    bpr = lg.LightGCL.calc_bpr_loss(me, local, users, pos, neg)
    g_bpr, = torch.autograd.grad(bpr, local)
    ssl = lg.LightGCL.calc_ssl_loss(me, local, glob, users, pos)
    g_ssl = torch.autograd.grad(ssl, [local, glob])
    emb_u = torch.nn.Embedding(n_users, D)
    emb_i = torch.nn.Embedding(n_items, D)
    me2 = SimpleNamespace(embedding_user=emb_u, embedding_item=emb_i)
    reg = lg.LightGCL.get_l2_reg(me2, users, pos - n_users, neg - n_users)
    g_reg = torch.autograd.grad(reg, [emb_u.weight, emb_i.weight])
    out = dict(local=local.detach(), glob=glob.detach(), users=users, pos=pos, neg=neg, temp=0.2, n_users=n_users,
               bpr=dict(loss=bpr.detach(), grad=g_bpr), ssl=dict(loss=ssl.detach(), grads=[x for x in g_ssl]),
               reg=dict(loss=reg.detach(), grads=[x for x in g_reg], user_w=emb_u.weight.detach(),
                        item_w=emb_i.weight.detach()))
    # retrieval: pure dot product, item 0 masked (v1_evaluate_lightgcl.py:312-318)
    all_items = torch.randn(n_items, D, generator=g)
    user_emb = torch.randn(40, D, generator=g)
    ns = dict(torch=torch, user_emb=user_emb, all_items=all_items, max_k=20)
    exec(ref_lines("gnn_model/v1_evaluate_lightgcl.py", 312, 318), ns)
    out["retrieval"] = dict(user_emb=user_emb, all_items=all_items, ids=ns["topk_indices"],
                            scores=torch.topk(ns["scores"], k=20, dim=1).values)
    save("lightgcl.pt", out)


if __name__ == "__main__":
    torch.set_num_threads(4)
    ru, ut, mi = import_tower_code()
    it = import_item_tower()
    make_user_tower(ru, ut)
    make_losses(ru, mi, it)
    make_retrieval()
    make_item_front(it)
    make_hybrid_user(mi)
    make_alignment(ut)
    make_ensemble()
    make_lightgcl()
