"""mmdet.apis.single_gpu_test / multi_gpu_test (mmdet/apis/test.py): iterate the loader, call
model(return_loss=False, rescale=True, **data), extend the result list."""
import torch


def single_gpu_test(model, data_loader, show=False, out_dir=None, show_score_thr=0.3):
    if show or out_dir:
        raise NotImplementedError('compat single_gpu_test: visualisation is outside the B200 backend')
    model.eval()
    results = []
    for data in data_loader:
        with torch.no_grad():
            result = model(return_loss=False, rescale=True, **data)
        results.extend(result)
    return results


def multi_gpu_test(model, data_loader, tmpdir=None, gpu_collect=False):
    """One rank per GPU; every rank runs its shard of the loader, results are gathered on rank 0 in dataset order."""
    import torch.distributed as dist
    results = single_gpu_test(model, data_loader)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return results
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, results)
    if dist.get_rank() != 0:
        return None
    ordered = []
    for i in range(max(len(p) for p in parts)):
        for p in parts:
            if i < len(p):
                ordered.append(p[i])
    return ordered[:len(data_loader.dataset)]
