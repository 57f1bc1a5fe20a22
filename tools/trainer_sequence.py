"""The reference trainer's step, call for call (trainer/trainer_egoclip.py:104-162), driving the package through the
reference-facing API only: `model(data)`, `sim_matrix`, `EgoNCE(x, sim_v, sim_n)`, an HF-style optimizer.

Test / benchmark infrastructure (tests/test_distributed_gpu.py, bench.py's `trainer_sequence` leg): the real trainer file
is unchanged and keeps its own `AllGather_multi`; this module restates it so that the "unchanged trainer" path -- four
list-API `dist.all_gather` calls, three `sim_matrix` launches, blocking fp32 `.to(device)` copies, `.item()` twice per
step -- can be measured and checked against the fused path (`egovlp_b200.distributed.egoclip_step_loss`) on the GPU box,
where /root/reference does not exist."""
import types

import torch
import torch.distributed as dist


class AllGatherMulti(torch.autograd.Function):
    """trainer/trainer_egoclip.py:11-27 (`AllGather_multi`): list-API all_gather + cat; backward = this rank's slice."""

    @staticmethod
    def forward(ctx, tensor, n_gpu, args):
        output = [torch.empty_like(tensor) for _ in range(args.world_size)]
        if args.world_size > 1:
            dist.all_gather(output, tensor)
        else:
            output[0].copy_(tensor)
        ctx.rank, ctx.batch_size = args.rank, tensor.shape[0]
        return torch.cat(output, 0)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output[ctx.batch_size * ctx.rank: ctx.batch_size * (ctx.rank + 1)], None, None


def dist_args():
    on = dist.is_available() and dist.is_initialized()
    return types.SimpleNamespace(world_size=dist.get_world_size() if on else 1, rank=dist.get_rank() if on else 0)


def trainer_step(model, loss_fn, optimizer, host_batch, device, sim_matrix, args=None, n_gpu=1):
    """One iteration of the loop body at trainer/trainer_egoclip.py:118-160 on an already tokenised host batch
    {'video', 'text': {'input_ids', 'attention_mask'}, 'noun_vec', 'verb_vec'}.  Returns the python float the trainer
    accumulates (`loss.detach().item()`, read twice as the trainer does with a writer attached)."""
    args = args or dist_args()
    data = dict(host_batch)
    data['text'] = {key: val.to(device) for key, val in data['text'].items()}            # :118 (blocking copies)
    data['video'] = data['video'].to(device)                                             # :119
    n_embeds = data['noun_vec'].to(device)                                               # :120
    v_embeds = data['verb_vec'].to(device)                                               # :121
    optimizer.zero_grad()                                                                # :123
    with torch.set_grad_enabled(True):
        text_embeds, video_embeds = model(data)                                          # :125
        video_embeds = AllGatherMulti.apply(video_embeds, n_gpu, args)                   # :126
        text_embeds = AllGatherMulti.apply(text_embeds, n_gpu, args)                     # :127
        n_embeds = AllGatherMulti.apply(n_embeds, n_gpu, args)                           # :128
        v_embeds = AllGatherMulti.apply(v_embeds, n_gpu, args)                           # :129
        output = sim_matrix(text_embeds, video_embeds)                                   # :130
        sim_v = sim_matrix(v_embeds, v_embeds)                                           # :133
        sim_n = sim_matrix(n_embeds, n_embeds)                                           # :134
        loss = loss_fn(output, sim_v, sim_n)                                             # :135
    loss.backward()                                                                      # :139
    optimizer.step()                                                                     # :141
    logged = loss.detach().item()                                                        # :148 (writer)
    total = loss.detach().item()                                                         # :150
    optimizer.zero_grad()                                                                # :160
    return total if logged == total else float("nan")
