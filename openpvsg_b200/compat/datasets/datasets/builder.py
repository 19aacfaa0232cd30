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

    def __init__(self, num_frames=8, height=96, width=160, seed=0, ref_seq_len_test=1, test_mode=True, ref_seq_index=None,
                 num_gt=3, **ignored):
        self.num_frames, self.hw, self.seed, self.T = int(num_frames), (int(height), int(width)), int(seed), int(ref_seq_len_test)
        self.test_mode = bool(test_mode)
        self.num_gt = int(num_gt)
        if not self.test_mode:                 # training clips: len(ref_seq_index) frames (2 in the reference config)
            self.T = len(ref_seq_index) if ref_seq_index else 2

    def __len__(self):
        return self.num_frames

    def _train_item(self, i):
        """One training clip in the format of the train pipeline's output (pvsg_vps.py:9-22): the ref_* ground truth of
        ``Mask2FormerVideoCustom.forward_train``."""
        h, w = self.hw
        d = syn.training_batch(1, h, w, self.T, self.num_gt, seed=self.seed + 13 * i)
        return dict(img=d['img'][0], img_metas=d['img_metas'][0], ref_img=d['ref_img'][0], ref_img_metas=d['ref_img_metas'][0],
                    ref_gt_bboxes=None, ref_gt_labels=d['ref_gt_labels'][0], ref_gt_masks=d['ref_gt_masks'][0],
                    ref_gt_semantic_seg=None, ref_gt_instance_ids=d['ref_gt_instance_ids'][0])

    def __getitem__(self, i):
        if not self.test_mode:
            return self._train_item(i)
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
