"""Continuous-batching generation engine over a paged KV cache.

Role parity: the serving stack the reference assembles from `block_multihead_attention` (paged KV cache + block tables,
paddle/phi/kernels/fusion/gpu/block_multi_head_attention_kernel.cu) and its in-flight batching scheduler: requests join and leave the
running batch between steps, the KV cache is a pool of fixed-size blocks handed out on demand, and a sequence that cannot get a block is
preempted (its blocks go back to the pool; it is re-prefilled later from prompt + generated tokens).

Every step packs the tokens of all scheduled sequences into ONE [tokens, hidden] batch: whole prompts for the sequences being prefilled,
one token for the sequences being decoded.  The decoder layers run on the packed batch with the model's own sublayers; attention is
`incubate.nn.paged_attention.block_attention` - on CUDA (head_dim 128, fp16 / bf16) one indexed scatter of the new K / V rows into the
block pool, `decode_attention_paged` for the decode rows and one packed variable-length tcgen05 attention for the prefill rows.
"""
from __future__ import annotations

import torch

from ..incubate.nn.paged_attention import block_attention
from .generation import _raw, _sample, _w


class BlockAllocator:
    """Pool of KV-cache blocks (ids 0 .. num_blocks-1)."""

    def __init__(self, num_blocks):
        self.num_blocks = int(num_blocks)
        self._free = list(range(self.num_blocks - 1, -1, -1))

    def num_free(self):
        return len(self._free)

    def alloc(self):
        if not self._free:
            raise MemoryError("KV-cache block pool is exhausted")
        return self._free.pop()

    def free(self, blocks):
        self._free.extend(reversed(blocks))


class Sequence:
    WAITING, RUNNING, FINISHED = "waiting", "running", "finished"

    def __init__(self, seq_id, prompt, max_new_tokens, eos_token_id=None, do_sample=False, temperature=1.0, top_k=0, top_p=1.0):
        self.id, self.prompt, self.max_new_tokens, self.eos = seq_id, [int(t) for t in prompt], int(max_new_tokens), eos_token_id
        self.do_sample, self.temperature, self.top_k, self.top_p = do_sample, temperature, top_k, top_p
        self.generated, self.blocks, self.cached, self.status, self.preemptions = [], [], 0, Sequence.WAITING, 0

    def tokens(self):
        return self.prompt + self.generated

    def finished(self):
        return self.status == Sequence.FINISHED


class LLMEngine:
    """add_request() any time; step() runs one scheduler iteration (admit / preempt, one packed forward, one token per running sequence)."""

    def __init__(self, model, num_blocks=256, block_size=16, max_running=64, max_batch_tokens=8192):
        from .generation import make_adapter

        model.eval()
        ad = make_adapter(model)                   # Llama (dense), Mixtral (MoE) and GPT layouts
        self.ad, self.model, self.cfg, self.layers = ad, model, ad.cfg, ad.layers
        self.nh, self.nkv, self.hd = ad.nh, ad.nkv, ad.hd
        self.block_size, self.max_running, self.max_batch_tokens = int(block_size), int(max_running), int(max_batch_tokens)
        p0 = _raw(next(iter(model.parameters())))
        self.device, self.dtype = p0.device, p0.dtype
        self.alloc = BlockAllocator(num_blocks)
        shape = (num_blocks, self.nkv, self.block_size, self.hd)
        self.key_cache = [torch.zeros(shape, dtype=self.dtype, device=self.device) for _ in self.layers]
        self.value_cache = [torch.zeros(shape, dtype=self.dtype, device=self.device) for _ in self.layers]
        self.waiting, self.running, self.done = [], [], {}
        self._next_id = 0
        self.stats = {"steps": 0, "prefill_tokens": 0, "decode_tokens": 0, "preemptions": 0, "max_running": 0}

    # ---- requests ---------------------------------------------------------------------------------------------------------------
    def add_request(self, prompt_ids, max_new_tokens=32, eos_token_id=None, do_sample=False, temperature=1.0, top_k=0, top_p=1.0):
        prompt = _raw(prompt_ids).reshape(-1).tolist() if isinstance(prompt_ids, torch.Tensor) else list(prompt_ids)
        need = (len(prompt) + max_new_tokens + self.block_size - 1) // self.block_size
        if need > self.alloc.num_blocks:
            raise ValueError(f"request needs {need} KV blocks, the pool has {self.alloc.num_blocks}")
        s = Sequence(self._next_id, prompt, max_new_tokens, eos_token_id, do_sample, temperature, top_k, top_p)
        self._next_id += 1
        self.waiting.append(s)
        return s.id

    def has_unfinished(self):
        return bool(self.waiting or self.running)

    # ---- scheduling -------------------------------------------------------------------------------------------------------------
    def _blocks_for(self, n_tokens):
        return (n_tokens + self.block_size - 1) // self.block_size

    def _preempt(self, s):
        self.alloc.free(s.blocks)
        s.blocks, s.cached, s.status = [], 0, Sequence.WAITING
        s.preemptions += 1
        self.stats["preemptions"] += 1
        self.running.remove(s)
        self.waiting.insert(0, s)              # first in line when blocks come back

    def _schedule(self):
        """Returns (decode sequences, prefill sequences).  Running sequences get their next slot first (newest ones are preempted when the
        pool runs dry); then waiting sequences are admitted while blocks, the running limit and the token budget allow."""
        decode = []
        for s in list(self.running):
            if s not in self.running:
                continue
            if self._blocks_for(s.cached + 1) > len(s.blocks):
                while self.alloc.num_free() == 0:
                    victim = next((v for v in reversed(self.running) if v is not s), None)
                    if victim is None:
                        break
                    self._preempt(victim)
                    if victim in decode:
                        decode.remove(victim)
                if self.alloc.num_free() == 0:
                    self._preempt(s)
                    continue
                s.blocks.append(self.alloc.alloc())
            decode.append(s)
        budget = self.max_batch_tokens - len(decode)
        prefill = []
        while self.waiting and len(self.running) + len(prefill) < self.max_running:
            s = self.waiting[0]
            n = len(s.tokens())
            need = self._blocks_for(n + 1)
            if n > budget or need > self.alloc.num_free():
                break
            self.waiting.pop(0)
            s.blocks = [self.alloc.alloc() for _ in range(need)]
            prefill.append(s)
            budget -= n
        return decode, prefill

    # ---- one packed forward -------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _forward(self, seqs, n_new, enc, dec):
        dev = self.device
        toks, pos = [], []
        for s, n in zip(seqs, n_new):
            all_t = s.tokens()
            toks += all_t[len(all_t) - n:]
            pos += list(range(s.cached, s.cached + n))
        ids = torch.tensor(toks, dtype=torch.int64, device=dev).unsqueeze(0)                 # [1, T]
        position_ids = torch.tensor(pos, dtype=torch.int64, device=dev).unsqueeze(0)
        cu = torch.zeros(len(seqs) + 1, dtype=torch.int32, device=dev)
        cu[1:] = torch.cumsum(torch.tensor(n_new, dtype=torch.int32, device=dev), 0)
        max_blocks = max(len(s.blocks) for s in seqs)
        bt = torch.zeros(len(seqs), max_blocks, dtype=torch.int32, device=dev)
        for i, s in enumerate(seqs):
            bt[i, : len(s.blocks)] = torch.tensor(s.blocks, dtype=torch.int32, device=dev)
        enc_t = torch.tensor(enc, dtype=torch.int32, device=dev)
        dec_t = torch.tensor(dec, dtype=torch.int32, device=dev)
        now_t = torch.tensor(n_new, dtype=torch.int32, device=dev)
        nh, nkv, hd = self.nh, self.nkv, self.hd
        h = self.ad.embed_tokens(ids, position_ids)
        for li, layer in enumerate(self.layers):
            qkv = self.ad.attn_in(layer, h, position_ids)
            t = qkv.shape[1]
            out, _, _, _ = block_attention(qkv.reshape(t, (nh + 2 * nkv) * hd), self.key_cache[li], self.value_cache[li], enc_t, dec_t, now_t, cu, bt, self.block_size)
            h = self.ad.attn_out(layer, h, _raw(out).reshape(1, t, nh * hd))
        last = (cu[1:].long() - 1)
        return self.ad.logits(_raw(h)[:, last])[0]                                     # [num_seqs, vocab]

    def step(self):
        """One iteration.  Returns [(request id, new token, finished)] for every sequence that produced a token."""
        decode, prefill = self._schedule()
        seqs = decode + prefill
        if not seqs:
            if self.waiting and not self.running:
                raise MemoryError("the KV-cache pool cannot hold the next waiting request")
            return []
        n_new = [1] * len(decode) + [len(s.tokens()) for s in prefill]
        enc = [0] * len(decode) + [len(s.tokens()) for s in prefill]
        dec = [s.cached for s in decode] + [0] * len(prefill)
        logits = self._forward(seqs, n_new, enc, dec)
        self.stats["steps"] += 1
        self.stats["decode_tokens"] += len(decode)
        self.stats["prefill_tokens"] += sum(n_new[len(decode):])
        out = []
        for i, (s, n) in enumerate(zip(seqs, n_new)):
            s.cached += n
            if s.status != Sequence.RUNNING:
                s.status = Sequence.RUNNING
                self.running.append(s)
            tok = int(_sample(logits[i: i + 1], s.do_sample, s.temperature, s.top_k, s.top_p)[0])
            s.generated.append(tok)
            fin = len(s.generated) >= s.max_new_tokens or (s.eos is not None and tok == s.eos)
            if fin:
                s.status = Sequence.FINISHED
                self.running.remove(s)
                self.alloc.free(s.blocks)
                s.blocks = []
                self.done[s.id] = s
            out.append((s.id, tok, fin))
        self.stats["max_running"] = max(self.stats["max_running"], len(self.running))
        return out

    def run_until_done(self, max_steps=100000):
        for _ in range(max_steps):
            if not self.has_unfinished():
                break
            self.step()
        return {i: s.generated for i, s in self.done.items()}

    def result(self, request_id):
        s = self.done.get(request_id)
        return None if s is None else _w(torch.tensor(s.generated, dtype=torch.int64))
