"""LineVul + DeepDFA combined model: the caller on the far side of the hot path (BASELINE configs[4]).

Host-side mirror of ``LineVul/linevul/linevul_model.py``: ``RobertaClassificationHead`` (:6-24 — ``cat(<s> feature, flowgnn
embedding) -> dropout -> dense -> tanh -> dropout -> out_proj(2)``) and ``Model.forward`` (:37-69 — same keyword signature, same
return tuples, ``CrossEntropyLoss``, ``softmax`` probabilities).  Submodule names are the reference's (``encoder``,
``flowgnn_encoder``, ``classifier.dense`` / ``classifier.out_proj``), so a checkpoint written by ``linevul_main.py`` loads.

What is new here is only scheduling: with ``overlap=True`` the DDFA encoder (``FlowGNNGGNNModule(encoder_mode=True)``,
hand-written kernels) is enqueued on a SIDE CUDA stream, so its launches overlap the transformer's forward on the main stream and
the two meet at the classifier head.  The transformer itself is the stock Hugging Face RoBERTa — out of scope of this repository.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import CrossEntropyLoss


class RobertaClassificationHead(nn.Module):
    """linevul_model.py:6-24."""

    def __init__(self, config, extra_dim):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size + extra_dim, config.hidden_size)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.out_proj = nn.Linear(config.hidden_size, 2)

    def forward(self, features, flowgnn_embed, **kwargs):
        x = features[:, 0, :]  # <s> token
        if flowgnn_embed is not None:
            x = torch.cat((x, flowgnn_embed), dim=1)
        x = self.dropout(x)
        x = self.dense(x)
        x = torch.tanh(x)
        x = self.dropout(x)
        return self.out_proj(x)


class LineVulCombined(nn.Module):
    """Drop-in for ``linevul_model.Model`` (linevul_model.py:26-69).  ``encoder`` is a ``RobertaForSequenceClassification``
    (only ``encoder.roberta`` is called, as in the reference), ``flowgnn_encoder`` any module with ``out_dim`` whose
    ``forward(graphs, {})`` returns ``[B, out_dim]`` — here the CUDA ``FlowGNNGGNNModule(encoder_mode=True)``.
    ``overlap=True`` enqueues the DDFA encoder on a side CUDA stream so that it runs concurrently with the transformer's forward
    (SURVEY.md §8 f3); default off: both are launch-bound from one host thread at the reference's batch size (16), and the side
    stream's event traffic cost more than the overlap bought in every measurement taken (tests/test_linevul.py prints both)."""

    def __init__(self, encoder, flowgnn_encoder, config, tokenizer=None, args=None, overlap: bool = False):
        super().__init__()
        self.encoder = encoder
        self.no_flowgnn = bool(getattr(args, "no_flowgnn", False)) if args is not None else flowgnn_encoder is None
        if not self.no_flowgnn:
            self.flowgnn_encoder = flowgnn_encoder
        self.tokenizer = tokenizer
        self.classifier = RobertaClassificationHead(config, 0 if self.no_flowgnn else self.flowgnn_encoder.out_dim)
        self.args = args
        self.overlap = overlap
        self._side = None

    def _flow_embed(self, graphs):
        """The DDFA embedding, enqueued on the side stream when the encoder lives on a CUDA device."""
        dev = next(self.flowgnn_encoder.parameters()).device
        if not self.overlap or dev.type != "cuda":
            return self.flowgnn_encoder(graphs, {}), None
        with torch.cuda.device(dev):
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev)
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)               # parameters / inputs produced on the main stream are ready
            with torch.cuda.stream(self._side):
                emb = self.flowgnn_encoder(graphs, {})
        return emb, self._side

    def forward(self, input_embed=None, labels=None, graphs=None, output_attentions=False, input_ids=None):
        flowgnn_embed, side = None, None
        if not self.no_flowgnn and graphs is not None:
            flowgnn_embed, side = self._flow_embed(graphs)          # runs concurrently with the transformer below
        if input_ids is not None:
            outputs = self.encoder.roberta(input_ids, attention_mask=input_ids.ne(1), output_attentions=output_attentions)
        else:
            outputs = self.encoder.roberta(inputs_embeds=input_embed, output_attentions=output_attentions)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)          # join: the head reads both
            flowgnn_embed.record_stream(torch.cuda.current_stream())
        last_hidden_state = outputs.last_hidden_state if output_attentions else outputs[0]
        logits = self.classifier(last_hidden_state, flowgnn_embed)
        prob = torch.softmax(logits, dim=-1)
        if labels is not None:
            loss = CrossEntropyLoss()(logits, labels)
            return (loss, prob, outputs.attentions) if output_attentions else (loss, prob)
        return (prob, outputs.attentions) if output_attentions else prob


def evaluate(model: LineVulCombined, batches, threshold: float = 0.5):
    """The scoring rule of ``linevul_main.evaluate`` (linevul_main.py:253-310): ``prob[:, 1] > 0.5`` against the labels,
    recall / precision / F1 by scikit-learn.  ``batches`` yields ``(input_ids, labels, graphs)``."""
    import numpy as np
    from sklearn.metrics import f1_score, precision_score, recall_score
    model.eval()
    probs, ys, loss_sum, steps = [], [], 0.0, 0
    for input_ids, labels, graphs in batches:
        with torch.no_grad():
            loss, prob = model(input_ids=input_ids, labels=labels, graphs=graphs)
        loss_sum += float(loss.mean())
        steps += 1
        probs.append(prob.detach().cpu().numpy())
        ys.append(labels.detach().cpu().numpy())
    probs, ys = np.concatenate(probs, 0), np.concatenate(ys, 0)
    pred = probs[:, 1] > threshold
    return {"eval_recall": float(recall_score(ys, pred, zero_division=0)), "eval_precision": float(precision_score(ys, pred, zero_division=0)),
            "eval_f1": float(f1_score(ys, pred, zero_division=0)), "eval_threshold": threshold, "eval_loss": loss_sum / max(steps, 1),
            "probs": probs, "labels": ys}
