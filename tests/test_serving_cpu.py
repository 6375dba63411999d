"""models.serving.LLMEngine: continuous batching over the paged KV cache gives every request exactly the tokens it gets when generated alone;
requests that arrive mid-flight join the running batch; a small block pool forces preemption and recomputation."""
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


def _alone(m, prompt, n):
    out = models.generate(m, torch.tensor([prompt]), max_new_tokens=n).as_subclass(torch.Tensor)
    return out[0, len(prompt):].tolist()


def test_block_allocator():
    a = models.BlockAllocator(4)
    got = [a.alloc() for _ in range(4)]
    assert sorted(got) == [0, 1, 2, 3] and a.num_free() == 0
    with pytest.raises(MemoryError):
        a.alloc()
    a.free(got[:2])
    assert a.num_free() == 2 and a.alloc() in got[:2]


def test_continuous_batching_matches_isolated_generation():
    m, cfg = _model()
    g = torch.Generator().manual_seed(0)
    prompts = [torch.randint(1, cfg.vocab_size, (n,), generator=g).tolist() for n in (5, 17, 9, 3)]
    news = [6, 4, 8, 5]
    eng = models.LLMEngine(m, num_blocks=64, block_size=4)
    ids = [eng.add_request(prompts[0], news[0]), eng.add_request(prompts[1], news[1])]
    eng.step()                                                     # both prompts prefilled in one packed batch
    eng.step()
    ids.append(eng.add_request(prompts[2], news[2]))               # joins while the first two are decoding
    eng.step()
    ids.append(eng.add_request(prompts[3], news[3]))
    res = eng.run_until_done()
    for i, p, n in zip(ids, prompts, news):
        assert res[i] == _alone(m, p, n), (i, res[i])
    assert eng.alloc.num_free() == 64 and eng.stats["preemptions"] == 0 and eng.stats["max_running"] >= 3
    assert eng.stats["prefill_tokens"] == sum(len(p) for p in prompts)
    assert eng.result(ids[0]).tolist() == res[ids[0]]


def test_preemption_when_the_block_pool_runs_dry():
    m, cfg = _model(seed=1)
    g = torch.Generator().manual_seed(1)
    prompts = [torch.randint(1, cfg.vocab_size, (n,), generator=g).tolist() for n in (6, 6, 6)]
    eng = models.LLMEngine(m, num_blocks=7, block_size=4)           # 3 x (6 + 10) tokens need 12 blocks: not everybody fits at once
    ids = [eng.add_request(p, 10) for p in prompts]
    res = eng.run_until_done()
    assert eng.stats["preemptions"] >= 1
    for i, p in zip(ids, prompts):
        assert res[i] == _alone(m, p, 10)
    assert eng.alloc.num_free() == 7
    with pytest.raises(ValueError):
        eng.add_request(list(range(1, 40)), 10)                     # can never fit


def test_eos_and_gqa_and_sampling_reproducibility():
    m, cfg = _model(seed=2, num_key_value_heads=2)
    p = [3, 14, 15, 92, 65]
    alone = _alone(m, p, 8)
    eng = models.LLMEngine(m, num_blocks=32, block_size=8)
    i = eng.add_request(p, 8, eos_token_id=alone[3])
    res = eng.run_until_done()
    assert res[i] == alone[:4]                                      # stops at (and includes) the EOS token
    paddle.seed(5)
    e1 = models.LLMEngine(m, num_blocks=32, block_size=8)
    a = e1.add_request(p, 6, do_sample=True, temperature=0.9, top_k=20)
    r1 = e1.run_until_done()[a]
    paddle.seed(5)
    e2 = models.LLMEngine(m, num_blocks=32, block_size=8)
    b = e2.add_request(p, 6, do_sample=True, temperature=0.9, top_k=20)
    assert e2.run_until_done()[b] == r1


def test_gpt_layout_through_the_engine():
    """Learned positions / LayerNorm / tied head: continuous batching reproduces per-request greedy decoding."""
    paddle.seed(6)
    cfg = models.gpt_tiny()
    m = models.GPTForCausalLM(cfg)
    m.eval()
    prompts = [torch.randint(1, cfg.vocab_size, (n,)).tolist() for n in (5, 9, 3)]
    eng = models.LLMEngine(m, num_blocks=24, block_size=4, max_running=2)
    ids = [eng.add_request(p, 6) for p in prompts]
    res = eng.run_until_done()
    for i, p in zip(ids, prompts):
        ref = models.generate(m, torch.tensor([p]), max_new_tokens=6).as_subclass(torch.Tensor)[0, len(p):].tolist()
        assert res[i] == ref
    assert eng.stats["max_running"] <= 2
