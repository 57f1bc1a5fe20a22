"""EgoMCQ metric of the reference's model/metric.py (:218-234) with GPU scoring (ops.egomcq_score)."""
import torch

from .. import ops


def egomcq_predict(text_embeds, video_embeds):
    """text [Q, C], video [Q, K, C] -> (scores [Q, K], pred [Q]); cosine similarity, ties -> lowest index."""
    return ops.egomcq_score(text_embeds.float(), video_embeds.float())


def egomcq_accuracy_metrics(preds, labels, types):
    """Same contract as the reference: preds [Q, K] scores, labels [Q], types [Q]; per-type accuracy in %.
    The reference zips ["Intra-video", "Inter-video"] with the SORTED unique type ids (:220-222); kept as is."""
    metrics = {}
    type_list = torch.unique(types)
    group_list = ["Intra-video", "Inter-video"]
    for type_i, group_i in zip(type_list, group_list):
        correct = total = 0
        for pred, label, typ in zip(preds, labels, types):
            if typ == type_i:
                pred_ = torch.argmax(pred)
                if pred_.item() == label.item():
                    correct += 1
                total += 1
        metrics[group_i] = correct / total * 100
    return metrics
