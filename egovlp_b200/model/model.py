"""FrozenInTime dual encoder + sim_matrix on the B200 kernels.

API mirror of the reference's model/model.py: FrozenInTime(video_params, text_params, projection_dim,
load_checkpoint, projection, load_temporal_fix), forward(data, video_only, return_embeds), compute_text,
compute_text_tokens, compute_video, set_device, sim_matrix(a, b, eps) -- and identical state_dict keys
(SURVEY.md section 8b), so reference checkpoints load and the unchanged trainer drives it.
"""
import os
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import engine
from .video_transformer import SpaceTimeTransformer


def state_dict_data_parallel_fix(load_state_dict, curr_state_dict):
    """Strip / add the DataParallel 'module.' prefix so that `load` matches `curr` (reference utils/util.py:25-51)."""
    load_keys, curr_keys = list(load_state_dict.keys()), list(curr_state_dict.keys())
    if not load_keys or not curr_keys:
        return load_state_dict
    have, want = load_keys[0].startswith('module.'), curr_keys[0].startswith('module.')
    if have and not want:
        return type(load_state_dict)((k[len('module.'):], v) for k, v in load_state_dict.items())
    if want and not have:
        return type(load_state_dict)(('module.' + k, v) for k, v in load_state_dict.items())
    return load_state_dict


class BaseModel(nn.Module):
    """Reference base/base_model.py: __str__ appends the trainable-parameter count."""

    def __str__(self):
        n = sum(int(np.prod(p.size())) for p in self.parameters() if p.requires_grad)
        return super().__str__() + '\nTrainable parameters: {}'.format(n)


def _build_distilbert(text_params):
    """Parameter container for the text tower: HuggingFace DistilBertModel (weights only; its forward is not used).
    Loads the pretrained files when present, otherwise random-initialises the same architecture."""
    from transformers import DistilBertConfig, DistilBertModel
    cache_dir = 'pretrained/distilbert-base-uncased'
    try:
        from transformers import AutoModel
        return AutoModel.from_pretrained('distilbert-base-uncased', cache_dir=cache_dir, local_files_only=True)
    except Exception:  # no network / no files: synthetic-weights path
        warnings.warn("distilbert-base-uncased files not found: text tower is randomly initialised")
        return DistilBertModel(DistilBertConfig())


_VIT_B16_FILE = "pretrained/jx_vit_base_p16_224-80ecf9dd.pth"


def _build_video_tower(video_params, from_scratch):
    """SpaceTimeTransformer exactly as the reference configures it (model/model.py:43-66): defaults for missing keys,
    ImageNet ViT-B/16 weights for a tower that is not restored from a checkpoint, identity head / pre_logits / fc."""
    kind = video_params['model']
    if kind != "SpaceTimeTransformer":
        raise NotImplementedError(f"{kind} not implemented")
    opts = {'num_frames': 4, 'time_init': 'zeros', 'attention_style': 'frozen-in-time', 'arch_config': 'base_patch16_224'}
    opts.update({k: video_params[k] for k in opts if k in video_params})
    if opts.pop('arch_config') != 'base_patch16_224':
        raise NotImplementedError
    tower = SpaceTimeTransformer(**opts)
    tower.head = tower.pre_logits = tower.fc = nn.Identity()      # `fc`: "backwards compatibility (old models)"
    if from_scratch:
        if os.path.exists(_VIT_B16_FILE):
            vit = torch.load(_VIT_B16_FILE, map_location="cpu")
            tower.load_state_dict(state_dict_data_parallel_fix(vit, tower.state_dict()), strict=False)
        else:
            warnings.warn(f"{_VIT_B16_FILE} not found: video tower keeps its random initialisation")
    return tower


class FrozenInTime(BaseModel):
    def __init__(self, video_params, text_params, projection_dim=256, load_checkpoint=None, projection='minimal',
                 load_temporal_fix='zeros'):
        super().__init__()
        self.video_params, self.text_params, self.load_temporal_fix = video_params, text_params, load_temporal_fix
        if not text_params['pretrained']:
            raise NotImplementedError("Huggingface text models require pretrained init.")
        if not text_params['model'].startswith('distilbert'):
            raise NotImplementedError(f"{text_params['model']}: only the DistilBERT text tower is implemented")
        self.text_model = _build_distilbert(text_params)
        self.text_model.train()                                   # as the reference: HF dropouts active while training
        restoring = load_checkpoint not in ("", None)
        self.video_model = _build_video_tower(video_params, from_scratch=not restoring)

        if projection == 'minimal':                               # project both towers to the common embedding
            self.txt_proj = nn.Sequential(nn.ReLU(), nn.Linear(self.text_model.config.hidden_size, projection_dim))
            self.vid_proj = nn.Sequential(nn.Linear(self.video_model.embed_dim, projection_dim))
        elif projection == '':
            self.txt_proj, self.vid_proj = nn.Identity(), nn.Identity()
        else:
            raise NotImplementedError
        object.__setattr__(self, "_bf16_cache", self.video_model._bf16_cache)
        if restoring:
            self._restore(load_checkpoint)

    def _restore(self, path):
        """model/model.py:88-95: a trainer checkpoint, saved with or without the DataParallel prefix and possibly with
        a different number of frames."""
        rank = int(os.environ.get('LOCAL_RANK', 0))
        where = f'cuda:{rank}' if torch.cuda.is_available() else 'cpu'
        saved = torch.load(path, map_location=where, weights_only=False)['state_dict']
        saved = self._inflate_positional_embeds(state_dict_data_parallel_fix(saved, self.state_dict()))
        self.load_state_dict(saved, strict=True)

    def set_device(self, device):
        self.device = device

    def forward(self, data, video_only=False, return_embeds=True):
        if video_only:
            return self.compute_video(data['video'])
        if torch.is_grad_enabled():
            self._bf16_cache.refresh()           # once per training forward (both towers share the cache)
        text_embeddings = self._text(data['text'], False, _refresh=False)
        video_embeddings = self.compute_video(data['video'], _refresh=False)
        if return_embeds:
            return text_embeddings, video_embeddings
        return sim_matrix(text_embeddings, video_embeddings)

    # ---- text ------------------------------------------------------------------------------------------------
    def _text_params(self):
        tm = self.text_model
        p = [tm.embeddings.word_embeddings.weight, tm.embeddings.position_embeddings.weight,
             tm.embeddings.LayerNorm.weight, tm.embeddings.LayerNorm.bias]
        for layer in tm.transformer.layer:
            a, f = layer.attention, layer.ffn
            for lin in (a.q_lin, a.k_lin, a.v_lin, a.out_lin):
                p += [lin.weight, lin.bias]
            p += [layer.sa_layer_norm.weight, layer.sa_layer_norm.bias, f.lin1.weight, f.lin1.bias, f.lin2.weight,
                  f.lin2.bias, layer.output_layer_norm.weight, layer.output_layer_norm.bias]
        return p

    def _text(self, text_data, tokens_mode, _refresh=True):
        # projection='' (nn.Identity, reference :80-82): the tower ends at the DistilBERT hidden state (no ReLU / Linear)
        proj = self.txt_proj[1] if isinstance(self.txt_proj, nn.Sequential) else None
        cfg = self.text_model.config
        if _refresh and torch.is_grad_enabled():
            self._bf16_cache.refresh()
        # train-mode dropouts of the HF text model (the reference keeps `text_model.train()`, :36); eval() or a config
        # with dropout = attention_dropout = 0 gives the deterministic path
        drop = (cfg.dropout, cfg.attention_dropout) if self.text_model.training else (0.0, 0.0)
        return engine.TextTowerFn.apply(text_data['input_ids'], text_data['attention_mask'], cfg.n_heads, 1e-12,
                                        (tokens_mode, torch.is_grad_enabled()), self._bf16_cache, drop, *self._text_params(),
                                        *((proj.weight, proj.bias) if proj is not None else (None, None)))

    def compute_text(self, text_data):
        return self._text(text_data, False)

    def compute_text_tokens(self, text_data):
        return self._text(text_data, True)

    # ---- video -----------------------------------------------------------------------------------------------
    def compute_video(self, video_data, _refresh=True):
        proj = self.vid_proj[0] if isinstance(self.vid_proj, nn.Sequential) else None
        return self.video_model.forward_features(video_data, proj=proj, _refresh=_refresh)

    # ---- checkpoint compat -----------------------------------------------------------------------------------
    def _inflate_positional_embeds(self, new_state_dict):
        """Load a checkpoint trained with a different number of frames (reference :145-187): truncate, or extend
        with zeros / nearest / bilinear interpolation along the frame axis."""
        key = 'video_model.temporal_embed'
        curr = self.state_dict()
        if key in new_state_dict and key in curr:
            load_te = new_state_dict[key]
            load_f, curr_f = load_te.shape[1], self.video_params['num_frames']
            if load_f > curr_f:
                new_state_dict[key] = load_te[:, :curr_f, :]
            elif load_f < curr_f:
                if self.load_temporal_fix == 'zeros':
                    new_te = torch.zeros([load_te.shape[0], curr_f, load_te.shape[2]], dtype=load_te.dtype,
                                         device=load_te.device)
                    new_te[:, :load_f] = load_te
                elif self.load_temporal_fix in ['interp', 'bilinear']:
                    mode = 'bilinear' if self.load_temporal_fix == 'bilinear' else 'nearest'
                    kw = dict(align_corners=True) if mode == 'bilinear' else {}
                    new_te = F.interpolate(load_te.unsqueeze(0), (curr_f, load_te.shape[2]), mode=mode, **kw).squeeze(0)
                else:
                    raise NotImplementedError
                new_state_dict[key] = new_te
        key = 'video_model.pos_embed'
        if key in new_state_dict and key in curr and new_state_dict[key].shape[1] != curr[key].shape[1]:
            raise NotImplementedError(
                'Loading models with different spatial resolution / patch number not yet implemented, sorry.')
        return new_state_dict


def sim_matrix(a, b, eps=1e-8):
    """Cosine similarity a_n @ b_n^T with norms clamped at eps (reference model/model.py:189-197)."""
    return engine.SimMatrixFn.apply(a, b, eps)
