"""Drop-in for the reference's model container (rnnt/model.py:6-139) whose loss path is the fused CUDA one."""
from __future__ import annotations

import torch

from .joint import JointNetwork
from .predictor import ConvPredictor, ConvPredictorStepper


class RNNTModel(torch.nn.Module):
    def __init__(self, predictor, encoder, joint):
        super().__init__()
        self.predictor = predictor
        self.encoder = encoder
        self.joint = joint

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, mel_features, mel_feature_lens, input_ids, input_id_lens, blank_idx: int):
        """Same contract as rnnt/model.py:17-43 (scalar mean loss); lines 32-41 collapse into joint.loss."""
        prepended = torch.cat([torch.full((input_ids.shape[0], 1), blank_idx, dtype=input_ids.dtype,
                                          device=input_ids.device), input_ids], dim=1)      # model.py:20-21
        decoder_features = self.predictor(prepended)                                         # model.py:24
        audio_features = self.encoder(mel_features).permute(0, 2, 1)                         # model.py:27-28
        audio_feature_lens = self.encoder.calc_output_lens(mel_feature_lens)                 # model.py:29
        return self.joint.loss(audio_features, decoder_features, input_ids.int(), audio_feature_lens.int(),
                               input_id_lens.int(), blank=-1, clamp=-1, reduction="mean")

    @torch.no_grad()
    def greedy_decode_features(self, audio_features, audio_feature_lens, max_length: int = 200,
                               max_outputs_per_step: int = 10, return_margins: bool = False,
                               use_cuda_graph: bool = True, sync_every: int = 16):
        """Batched greedy decode on given encoder features (B,T,H) with per-utterance lengths.

        Same per-utterance algorithm as rnnt/model.py:90-128 (blank or 10 emits advances the frame; stop at T_b or
        when len(tokens) incl. the seed blank reaches max_length), run for the whole batch at once with ALL loop state
        on the device: per step one joint+argmax kernel call for the batch, one incremental predictor update, and a
        handful of elementwise ops -- captured once in a CUDA graph and replayed; the host only checks for completion
        every `sync_every` steps (the reference syncs on `.item()` every step, model.py:113)."""
        if not isinstance(self.predictor, ConvPredictor):
            raise ValueError("batched greedy decode supports ConvPredictor")
        if not audio_features.is_cuda:
            raise RuntimeError("rnnt_b200 decode runs on CUDA tensors only; there is no CPU fallback")
        from .functional import joint_argmax_scratch
        dev = audio_features.device
        B, T, _ = audio_features.shape
        V = self.joint.joint_ln.weight.shape[0]
        blank = self.joint.blank_idx
        audio_features = audio_features.contiguous()
        lens = audio_feature_lens.to(dev, torch.int64).clamp(max=T)
        was_training = self.predictor.training
        self.predictor.eval()
        try:
            stepper = ConvPredictorStepper(self.predictor, B, dev)
            rows = torch.arange(B, device=dev)
            t_idx = torch.zeros(B, dtype=torch.int64, device=dev)
            per = torch.zeros(B, dtype=torch.int64, device=dev)
            ntok = torch.ones(B, dtype=torch.int64, device=dev)          # counts the seed blank (model.py:53,64)
            out = torch.full((B, max(max_length, 1)), blank, dtype=torch.int64, device=dev)
            all_rows = torch.ones(B, dtype=torch.bool, device=dev)
            feats = stepper.advance(torch.full((B,), blank, dtype=torch.int64, device=dev), all_rows).contiguous()
            tok32 = torch.empty(B, dtype=torch.int32, device=dev)
            margin = torch.empty(B, dtype=torch.float32, device=dev)
            scratch = joint_argmax_scratch(B, V, dev)
            max_steps = int(lens.max()) + max_length + 1
            margin_log = torch.full((max_steps, B), float("inf"), device=dev) if return_margins else None
            step_no = torch.zeros((), dtype=torch.int64, device=dev)
            inf_row = torch.full((B,), float("inf"), device=dev)

            def one_step():
                active = (t_idx < lens) & (ntok < max_length)
                a_rows = audio_features[rows, t_idx.clamp(max=T - 1)]
                self.joint.argmax_step(a_rows, feats, return_margin=True, out=tok32, margin_out=margin,
                                       scratch=scratch)
                tok = tok32.to(torch.int64)
                advance = active & ((tok == blank) | (per >= max_outputs_per_step))
                emit = active & ~advance
                if margin_log is not None:
                    margin_log[step_no.clamp(max=max_steps - 1)] = torch.where(active, margin, inf_row)
                    step_no.add_(1)
                t_idx.add_(advance.to(torch.int64))
                per.copy_(torch.where(advance, torch.zeros_like(per), per + emit.to(torch.int64)))
                pos = ntok.clamp(max=out.shape[1]) - 1                      # token k (0-based, seed excluded) -> column k
                cur = out[rows, pos.clamp(min=0)]
                out[rows, pos.clamp(min=0)] = torch.where(emit, tok, cur)
                ntok.add_(emit.to(torch.int64))
                feats.copy_(torch.where(emit.view(-1, 1), stepper.advance(tok, emit), feats))

            graph = None
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                if use_cuda_graph:
                    try:
                        saved = [x.clone() for x in (t_idx, per, ntok, out, feats, stepper.xs, stepper.ys, step_no)]
                        side = torch.cuda.Stream(device=dev)
                        side.wait_stream(torch.cuda.current_stream(dev))
                        with torch.cuda.stream(side):
                            one_step()                                  # warm-up outside capture
                        torch.cuda.current_stream(dev).wait_stream(side)
                        for x, sv in zip((t_idx, per, ntok, out, feats, stepper.xs, stepper.ys, step_no), saved):
                            x.copy_(sv)
                        if margin_log is not None:
                            margin_log.fill_(float("inf"))
                        graph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(graph):
                            one_step()
                        for x, sv in zip((t_idx, per, ntok, out, feats, stepper.xs, stepper.ys, step_no), saved):
                            x.copy_(sv)
                        if margin_log is not None:
                            margin_log.fill_(float("inf"))
                    except Exception:
                        graph = None
                        for x, sv in zip((t_idx, per, ntok, out, feats, stepper.xs, stepper.ys, step_no), saved):
                            x.copy_(sv)
                done_steps = 0
                while done_steps < max_steps:
                    for _ in range(min(sync_every, max_steps - done_steps)):
                        if graph is not None:
                            graph.replay()
                        else:
                            one_step()
                        done_steps += 1
                    if not bool(((t_idx < lens) & (ntok < max_length)).any()):
                        break
        finally:
            self.predictor.train(was_training)
        n = (ntok - 1).tolist()
        out_host = out.tolist()
        result = [out_host[b][: n[b]] for b in range(B)]
        if return_margins:
            ml = margin_log[:done_steps].t().tolist()
            margins = [[m for m in ml[b] if m != float("inf")] for b in range(B)]
            return result, margins
        return result

    @torch.no_grad()
    def _greedy_decode_features_hostloop(self, audio_features, audio_feature_lens, max_length: int = 200,
                                         max_outputs_per_step: int = 10):
        """Host-driven variant (one device sync per step, predictor re-run on 7-token windows); kept as an independent
        cross-check of the device-side loop in the tests."""
        dev = audio_features.device
        B, T, _ = audio_features.shape
        blank = self.joint.blank_idx
        lens = [int(x) for x in audio_feature_lens.tolist()]
        tokens = [[blank] for _ in range(B)]
        t_idx = [0] * B
        per = [0] * B
        was_training = self.predictor.training
        self.predictor.eval()
        try:
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                feats = self.predictor(torch.full((B, 1), blank, dtype=torch.int64, device=dev))[:, -1, :].contiguous()
                active = [b for b in range(B) if t_idx[b] < lens[b] and len(tokens[b]) < max_length]
                while active:
                    idx = torch.tensor(active, device=dev)
                    tt = torch.tensor([t_idx[b] for b in active], device=dev)
                    toks = self.joint.argmax_step(audio_features[idx, tt], feats[idx]).tolist()
                    emitted = []
                    for j, b in enumerate(active):
                        if toks[j] == blank or per[b] >= max_outputs_per_step:
                            t_idx[b] += 1
                            per[b] = 0
                        else:
                            tokens[b].append(toks[j])
                            per[b] += 1
                            emitted.append(b)
                    by_len = {}
                    for b in emitted:
                        by_len.setdefault(min(len(tokens[b]), ConvPredictor.RECEPTIVE_FIELD), []).append(b)
                    for wl, group in by_len.items():
                        win = torch.tensor([tokens[b][-wl:] for b in group], dtype=torch.int64, device=dev)
                        feats[torch.tensor(group, device=dev)] = self.predictor.last_step(win)
                    active = [b for b in active if t_idx[b] < lens[b] and len(tokens[b]) < max_length]
        finally:
            self.predictor.train(was_training)
        return [tk[1:] for tk in tokens]

    @torch.no_grad()
    def greedy_decode(self, mel_features, mel_feature_lens, max_length: int = 200):
        """Reference signature (rnnt/model.py:130-139): batch of one, returns list[int]."""
        assert mel_features.shape[0] == 1, "Greedy decoding only works with a batch size of 1"
        audio_features = self.encoder(mel_features).permute(0, 2, 1).contiguous()
        # the reference loops to audio_features.shape[1] (model.py:56,100), not to the computed length
        lens = torch.tensor([audio_features.shape[1]])
        return self.greedy_decode_features(audio_features, lens, max_length)[0]


def enable_zero_edit_mode(joint: JointNetwork, enabled: bool = True) -> None:
    """Make the UNMODIFIED reference rnnt/model.py:32-41 run fused: joint.forward hands a lazy handle to a patched
    torchaudio.functional.rnnt_loss.  Call once after building the model."""
    import torchaudio
    from . import functional
    joint.zero_edit_mode = enabled
    if enabled and getattr(torchaudio.functional, "_rnnt_b200_original_rnnt_loss", None) is None:
        original = torchaudio.functional.rnnt_loss

        def patched(logits, targets, logit_lengths, target_lengths, blank=-1, clamp=-1, reduction="mean",
                    fused_log_softmax=True):
            from .joint import LazyJointLogits
            if isinstance(logits, LazyJointLogits):
                return functional.rnnt_loss(logits, targets, logit_lengths, target_lengths, blank, clamp, reduction)
            return original(logits, targets, logit_lengths, target_lengths, blank, clamp, reduction,
                            fused_log_softmax)

        torchaudio.functional._rnnt_b200_original_rnnt_loss = original
        torchaudio.functional.rnnt_loss = patched
