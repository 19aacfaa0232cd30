"""datasets/datasets/builder.py::build_dataset with the dataset types this backend ships."""
import torch
from torch.utils.data import Dataset

from openpvsg_b200 import synthetic as syn
from openpvsg_b200.registry import Registry, build_from_cfg

DATASETS = Registry('dataset')
PIPELINES = Registry('pipeline')


@DATASETS.register_module()
class SyntheticVPSDataset(Dataset):
    """A clip of seeded synthetic frames in the sample format of the reference's VPS test pipeline (SeqNormalize,
    SeqPad(32), VideoCollect, ConcatVideoReferences, SeqDefaultFormatBundle; configs/_base_/datasets/pvsg_vps.py:27-34):
    img [3,Hp,Wp], img_metas, ref_img [T,3,Hp,Wp], ref_img_metas [T].  Unknown reference keys (split, video_name,
    pipeline, ref_sample_mode, ...) are accepted and ignored so `--cfg-options data.test.type=SyntheticVPSDataset`
    works on an unmodified reference config."""
    CLASSES = tuple(f'class_{i}' for i in range(126))

    def __init__(self, num_frames=8, height=96, width=160, seed=0, ref_seq_len_test=1, test_mode=True, **ignored):
        self.num_frames, self.hw, self.seed, self.T = int(num_frames), (int(height), int(width)), int(seed), int(ref_seq_len_test)

    def __len__(self):
        return self.num_frames

    def __getitem__(self, i):
        h, w = self.hw
        frames = torch.stack([syn.synthetic_frame(self.seed + i + t, h, w) for t in range(self.T)])
        meta = syn.frame_meta(h, w)
        meta.pop('batch_input_shape')          # forward_test adds it
        return dict(img=frames[0], img_metas=dict(meta), ref_img=frames, ref_img_metas=[dict(meta) for _ in range(self.T)])


def build_dataset(cfg, default_args=None):
    typ = cfg.get('type') if isinstance(cfg, dict) else None
    if isinstance(typ, str) and typ not in DATASETS:
        raise KeyError(f'dataset type {typ!r}: the disk-backed PVSG datasets / pipelines belong to the reference and need its '
                       'mmdet 2.25 stack; this backend ships SyntheticVPSDataset (--cfg-options data.test.type=SyntheticVPSDataset)')
    return build_from_cfg(dict(cfg), DATASETS, default_args)
