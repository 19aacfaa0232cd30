"""mmdet.datasets.build_dataloader / replace_ImageToTensor for test-mode loaders."""
import torch
from torch.utils.data import DataLoader, DistributedSampler


def _collate(batch):
    """Samples are dicts of tensors (stacked over the batch) and metas (kept as lists): the call signature of
    BaseDetector.forward_test / Mask2FormerVideoCustom.forward_test -- img=[B,3,H,W] wrapped in the test-time
    augmentation list, img_metas=[[dict] * B], ref_img=[B,T,3,H,W], ref_img_metas=[[dict] * T] * B."""
    out = {}
    for key in batch[0]:
        vals = [b[key] for b in batch]
        if key == 'img':
            out[key] = [torch.stack(vals)]
        elif key == 'img_metas':
            out[key] = [vals]
        elif torch.is_tensor(vals[0]):
            out[key] = torch.stack(vals)
        else:
            out[key] = vals
    return out


def _collate_train(batch):
    """Training samples (SeqDefaultFormatBundle + Collect of the train pipeline): image tensors are stacked (img [B,3,H,W],
    ref_img [B,T,3,H,W]); metas and every ground-truth field stay lists over the batch -- the keyword arguments of
    Mask2FormerVideoCustom.forward_train."""
    out = {}
    for key in batch[0]:
        vals = [b[key] for b in batch]
        out[key] = torch.stack(vals) if key in ('img', 'ref_img') else vals
    return out


def build_dataloader(dataset, samples_per_gpu=1, workers_per_gpu=0, num_gpus=1, dist=False, shuffle=False, seed=None,
                     persistent_workers=False, train=False, **kwargs):
    if train:
        gen = torch.Generator()
        gen.manual_seed(int(seed) if seed is not None else 0)
        if dist:
            return DataLoader(dataset, batch_size=samples_per_gpu, sampler=DistributedSampler(dataset, shuffle=shuffle, seed=int(seed or 0)),
                              num_workers=workers_per_gpu, collate_fn=_collate_train, drop_last=False)
        return DataLoader(dataset, batch_size=samples_per_gpu, shuffle=shuffle, generator=gen, num_workers=workers_per_gpu,
                          collate_fn=_collate_train, drop_last=False)
    sampler = DistributedSampler(dataset, shuffle=False) if dist else None
    return DataLoader(dataset, batch_size=samples_per_gpu, sampler=sampler, shuffle=False, num_workers=workers_per_gpu,
                      collate_fn=_collate, pin_memory=torch.cuda.is_available(),
                      persistent_workers=persistent_workers and workers_per_gpu > 0)


def replace_ImageToTensor(pipelines):
    return pipelines
