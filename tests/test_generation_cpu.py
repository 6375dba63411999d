"""models.generation: KV-cache generation equals recomputing the whole sequence every step; ragged prompts; sampling controls; EOS."""
import numpy as np
import pytest
import torch

import paddle_b200 as paddle
from paddle_b200 import models


def _model(seed=0, **kw):
    paddle.seed(seed)
    cfg = models.llama_tiny(**kw)
    m = models.LlamaForCausalLM(cfg)
    m.eval()
    return m, cfg


def _naive_greedy(m, ids, steps):
    ids = ids.clone()
    for _ in range(steps):
        with torch.no_grad():
            logits = m(ids.as_subclass(paddle.Tensor)).as_subclass(torch.Tensor)
        ids = torch.cat([ids, logits[:, -1].argmax(-1, keepdim=True)], 1)
    return ids


def test_greedy_generation_matches_full_recompute():
    m, cfg = _model()
    ids = torch.randint(0, cfg.vocab_size, (3, 7))
    ref = _naive_greedy(m, ids, 6)
    out, cache = models.generate(m, ids.as_subclass(paddle.Tensor), max_new_tokens=6, return_cache=True)
    assert torch.equal(out.as_subclass(torch.Tensor), ref)
    assert cache.lens.tolist() == [12, 12, 12] and cache.k[0].shape == (3, cfg.num_key_value_heads, 13, cfg.head_dim)
    assert cache.nbytes() == 2 * cfg.num_hidden_layers * 3 * cfg.num_key_value_heads * 13 * cfg.head_dim * 4


def test_grouped_query_attention_cache():
    m, cfg = _model(seed=1, num_key_value_heads=2)
    assert cfg.num_key_value_heads == 2 and cfg.num_attention_heads > 2
    ids = torch.randint(0, cfg.vocab_size, (2, 5))
    assert torch.equal(models.generate(m, ids, max_new_tokens=4).as_subclass(torch.Tensor), _naive_greedy(m, ids, 4))


def test_ragged_prompts_right_padded():
    m, cfg = _model(seed=2)
    a, b = torch.randint(1, cfg.vocab_size, (1, 9)), torch.randint(1, cfg.vocab_size, (1, 4))
    ref_a, ref_b = _naive_greedy(m, a, 5), _naive_greedy(m, b, 5)
    batch = torch.zeros(2, 9, dtype=torch.int64)
    batch[0], batch[1, :4] = a[0], b[0]
    out = models.generate(m, batch, max_new_tokens=5, prompt_lens=torch.tensor([9, 4]), pad_token_id=0).as_subclass(torch.Tensor)
    assert torch.equal(out[0, :14], ref_a[0])
    assert torch.equal(out[1, :9], ref_b[0])                    # the shorter prompt's continuation starts right after ITS last token
    assert out.shape == (2, 14)


def test_sampling_controls_and_eos():
    m, cfg = _model(seed=3)
    ids = torch.randint(0, cfg.vocab_size, (2, 6))
    greedy = models.generate(m, ids, max_new_tokens=5).as_subclass(torch.Tensor)
    # top_k = 1 sampling is greedy whatever the temperature; so is top_p -> 0
    assert torch.equal(models.generate(m, ids, max_new_tokens=5, do_sample=True, temperature=1.7, top_k=1).as_subclass(torch.Tensor), greedy)
    assert torch.equal(models.generate(m, ids, max_new_tokens=5, do_sample=True, top_p=1e-6).as_subclass(torch.Tensor), greedy)
    # sampling is reproducible under the framework seed and differs between seeds
    paddle.seed(7)
    s1 = models.generate(m, ids, max_new_tokens=8, do_sample=True, temperature=1.0).as_subclass(torch.Tensor)
    paddle.seed(7)
    s2 = models.generate(m, ids, max_new_tokens=8, do_sample=True, temperature=1.0).as_subclass(torch.Tensor)
    paddle.seed(8)
    s3 = models.generate(m, ids, max_new_tokens=8, do_sample=True, temperature=1.0).as_subclass(torch.Tensor)
    assert torch.equal(s1, s2) and not torch.equal(s1, s3)
    # EOS: once a sequence emits it, the rest is padding and generation can stop early
    eos = int(greedy[0, 6])
    out = models.generate(m, ids, max_new_tokens=5, eos_token_id=eos, pad_token_id=cfg.vocab_size - 1).as_subclass(torch.Tensor)
    assert int(out[0, 6]) == eos and (out[0, 7:] == cfg.vocab_size - 1).all()
    assert m.training is False
    m.train()
    models.generate(m, ids, max_new_tokens=1)
    assert m.training is True                                  # the caller's mode is restored


def test_mixtral_moe_generation_and_serving():
    """The MoE feed-forward routes per token, so cached decoding (and continuous batching) reproduces full recomputation."""
    paddle.seed(4)
    cfg = models.mixtral_tiny()
    m = models.MixtralForCausalLM(cfg)
    m.eval()
    ids = torch.randint(1, cfg.vocab_size, (2, 6))
    ref = ids.clone()
    for _ in range(5):
        with torch.no_grad():
            lg = m(ref.as_subclass(paddle.Tensor)).as_subclass(torch.Tensor)
        ref = torch.cat([ref, lg[:, -1].argmax(-1, keepdim=True)], 1)
    out = models.generate(m, ids, max_new_tokens=5).as_subclass(torch.Tensor)
    assert torch.equal(out, ref)
    eng = models.LLMEngine(m, num_blocks=32, block_size=4)
    a, b = eng.add_request(ids[0].tolist(), 5), eng.add_request(ids[1].tolist(), 5)
    res = eng.run_until_done()
    assert res[a] == ref[0, 6:].tolist() and res[b] == ref[1, 6:].tolist()


def test_gpt_generation_matches_full_recompute():
    """GPT layout: learned positions, LayerNorm, tied head.  Greedy KV-cache decoding equals recomputing the sequence; ragged prompts."""
    paddle.seed(5)
    cfg = models.gpt_tiny()
    m = models.GPTForCausalLM(cfg)
    m.eval()
    ids = torch.randint(1, cfg.vocab_size, (2, 7))
    ref = ids.clone()
    for _ in range(6):
        with torch.no_grad():
            lg = m(ref.as_subclass(paddle.Tensor)).as_subclass(torch.Tensor)
        ref = torch.cat([ref, lg[:, -1].argmax(-1, keepdim=True)], 1)
    assert torch.equal(models.generate(m, ids, max_new_tokens=6).as_subclass(torch.Tensor), ref)
    batch = ids.clone()
    batch[1, 4:] = 0
    out = models.generate(m, batch, max_new_tokens=4, prompt_lens=torch.tensor([7, 4])).as_subclass(torch.Tensor)
    solo = models.generate(m, ids[1:2, :4], max_new_tokens=4).as_subclass(torch.Tensor)
    assert torch.equal(out[1, :8], solo[0])
    with pytest.raises(ValueError):
        models.generate(m, ids, max_new_tokens=cfg.max_position_embeddings)
