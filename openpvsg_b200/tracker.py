"""IPS tracker path: UniTrack-style association of per-frame instance masks into tubes.

Reference: ``models/unitrack/multitracker.py:27-199`` (``class_aware_distance``, ``AssociationTracker.update``),
``models/unitrack/mask.py:17-60`` (``MaskAssociationTracker``), ``models/unitrack/basetrack.py`` (``STrack`` and the
list helpers), ``core/association/matching.py`` (``reconsdot_distance`` :194-238, ``linear_assignment`` :29-40,
``iou_distance`` :62-79, ``fuse_motion`` :98-110), ``core/motion/kalman_filter.py``, ``data/query_feat_tracklet.py``;
driven by ``tools/prepare_query_tube_ips.py:256-260`` -> ``test_mots_from_mask2former.py::eval_seq``.

Same class / function names, arguments and return values.  What runs where:

* device (libpvsg_sm100.so): the appearance distance ``reconsdot_distance`` (``pvsg_reconsdot``), the assignment
  ``linear_assignment`` (``pvsg_lap_assign``, replaces the absent ``lap`` wheel), the bilinear rescale of large masks'
  features in ``extract_emb``;
* host: the track bookkeeping (Python lists of a few dozen tracks, as the reference), the 8-state Kalman filter and the
  box IoU (numpy fp64 on 4-vectors, as the reference; ``cython_bbox`` restated).

The appearance network of the reference (``models/unitrack/model``, weights ``checkpoints/UniTrack/timecycle.pth``) is an
external asset: ``MaskAssociationTracker`` takes any callable ``app_model(img [1,3,H,W]) -> [1,d,h,w]``; by default the
detector's own backbone stage (stride 8) is used.
"""
from collections import deque

import numpy as np
import torch

from . import ops


# ======================================================================================
# core/motion/kalman_filter.py
# ======================================================================================
chi2inv95 = {1: 3.8415, 2: 5.9915, 3: 7.8147, 4: 9.4877, 5: 11.070, 6: 12.592, 7: 14.067, 8: 15.507, 9: 16.919}


class KalmanFilter:
    """Constant-velocity filter on (x, y, a, h, vx, vy, va, vh) -- kalman_filter.py:23-277."""

    def __init__(self):
        ndim, dt = 4, 1.
        self._motion_mat = np.eye(2 * ndim, 2 * ndim)
        for i in range(ndim):
            self._motion_mat[i, ndim + i] = dt
        self._update_mat = np.eye(ndim, 2 * ndim)
        self._std_weight_position = 1. / 20
        self._std_weight_velocity = 1. / 160

    def _stds(self, h, scale_pos=1.0, scale_vel=1.0):
        p, v = self._std_weight_position * h, self._std_weight_velocity * h
        return [scale_pos * p, scale_pos * p, 1e-2, scale_pos * p], [scale_vel * v, scale_vel * v, 1e-5, scale_vel * v]

    def initiate(self, measurement):
        mean = np.r_[measurement, np.zeros_like(measurement)]
        sp, sv = self._stds(measurement[3], 2.0, 10.0)
        return mean, np.diag(np.square(np.r_[sp, sv]))

    def predict(self, mean, covariance):
        sp, sv = self._stds(mean[3])
        motion_cov = np.diag(np.square(np.r_[sp, sv]))
        mean = np.dot(mean, self._motion_mat.T)
        covariance = np.linalg.multi_dot((self._motion_mat, covariance, self._motion_mat.T)) + motion_cov
        return mean, covariance

    def multi_predict(self, mean, covariance):
        h = mean[:, 3]
        std = np.stack([self._std_weight_position * h, self._std_weight_position * h, 1e-2 * np.ones_like(h),
                        self._std_weight_position * h, self._std_weight_velocity * h, self._std_weight_velocity * h,
                        1e-5 * np.ones_like(h), self._std_weight_velocity * h], 1)
        motion_cov = np.stack([np.diag(s) for s in np.square(std)])
        mean = np.dot(mean, self._motion_mat.T)
        left = np.dot(self._motion_mat, covariance).transpose((1, 0, 2))
        return mean, np.dot(left, self._motion_mat.T) + motion_cov

    def project(self, mean, covariance):
        std = [self._std_weight_position * mean[3], self._std_weight_position * mean[3], 1e-1,
               self._std_weight_position * mean[3]]
        mean = np.dot(self._update_mat, mean)
        covariance = np.linalg.multi_dot((self._update_mat, covariance, self._update_mat.T))
        return mean, covariance + np.diag(np.square(std))

    def update(self, mean, covariance, measurement):
        import scipy.linalg
        projected_mean, projected_cov = self.project(mean, covariance)
        chol, lower = scipy.linalg.cho_factor(projected_cov, lower=True, check_finite=False)
        gain = scipy.linalg.cho_solve((chol, lower), np.dot(covariance, self._update_mat.T).T, check_finite=False).T
        innovation = measurement - projected_mean
        return mean + np.dot(innovation, gain.T), covariance - np.linalg.multi_dot((gain, projected_cov, gain.T))

    def gating_distance(self, mean, covariance, measurements, only_position=False, metric='maha'):
        import scipy.linalg
        mean, covariance = self.project(mean, covariance)
        if only_position:
            mean, covariance, measurements = mean[:2], covariance[:2, :2], measurements[:, :2]
        d = measurements - mean
        if metric == 'gaussian':
            return np.sum(d * d, axis=1)
        if metric == 'maha':
            z = scipy.linalg.solve_triangular(np.linalg.cholesky(covariance), d.T, lower=True, check_finite=False,
                                              overwrite_b=True)
            return np.sum(z * z, axis=0)
        raise ValueError('invalid distance metric')


# ======================================================================================
# utils/box.py (the conversions the tracker uses), basetrack.py
# ======================================================================================
def tlwh_to_xyah(tlwh):
    """utils/box.py:18-25 (note the 1e-6 in the aspect ratio)."""
    ret = np.asarray(tlwh).copy()
    ret[:2] += ret[2:] / 2
    ret[2] /= (ret[3] + 1e-6)
    return ret


def tlbr_to_tlwh(tlbr):
    ret = np.asarray(tlbr).copy()
    ret[2:] -= ret[:2]
    return ret


class TrackState:
    New, Tracked, Lost, Removed = 0, 1, 2, 3


class BaseTrack:
    _count = 0
    track_id = 0
    is_activated = False
    state = TrackState.New
    score = 0
    start_frame = 0
    frame_id = 0

    @property
    def end_frame(self):
        return self.frame_id

    @staticmethod
    def next_id():
        BaseTrack._count += 1
        return BaseTrack._count

    @staticmethod
    def reset_count():
        BaseTrack._count = 0

    def mark_lost(self):
        self.state = TrackState.Lost

    def mark_removed(self):
        self.state = TrackState.Removed


class STrack(BaseTrack):
    """basetrack.py:63-231."""
    shared_kalman = KalmanFilter()

    def __init__(self, tlwh, score, temp_feat, buffer_size=30, mask=None, pose=None, ac=False, category=-1, use_kalman=True):
        self._tlwh = np.asarray(tlwh, dtype=np.float64)
        self.kalman_filter = None
        self.mean, self.covariance = None, None
        self.use_kalman = use_kalman
        if not use_kalman:
            ac = True
        self.is_activated = ac
        self.score = score
        self.category = category
        self.tracklet_len = 0
        self.smooth_feat = None
        self.update_features(temp_feat)
        self.features = deque([], maxlen=buffer_size)
        self.alpha = 0.9
        self.mask = mask
        self.pose = pose

    def update_features(self, feat):
        self.curr_feat = feat
        if self.smooth_feat is None:
            self.smooth_feat = feat
        elif self.smooth_feat.shape == feat.shape:
            self.smooth_feat = self.alpha * self.smooth_feat + (1 - self.alpha) * feat

    def predict(self):
        mean_state = self.mean.copy()
        if self.state != TrackState.Tracked:
            mean_state[7] = 0
        self.mean, self.covariance = self.kalman_filter.predict(mean_state, self.covariance)

    @staticmethod
    def multi_predict(stracks):
        if len(stracks) > 0:
            multi_mean = np.asarray([st.mean.copy() for st in stracks])
            multi_covariance = np.asarray([st.covariance for st in stracks])
            for i, st in enumerate(stracks):
                if st.state != TrackState.Tracked:
                    multi_mean[i][7] = 0
            multi_mean, multi_covariance = STrack.shared_kalman.multi_predict(multi_mean, multi_covariance)
            for i, (mean, cov) in enumerate(zip(multi_mean, multi_covariance)):
                stracks[i].mean, stracks[i].covariance = mean, cov

    def activate(self, kalman_filter, frame_id):
        self.kalman_filter = kalman_filter
        self.track_id = self.next_id()
        self.mean, self.covariance = self.kalman_filter.initiate(tlwh_to_xyah(self._tlwh))
        self.tracklet_len = 0
        self.state = TrackState.Tracked
        if frame_id == 1:
            self.is_activated = True
        self.frame_id = frame_id
        self.start_frame = frame_id

    def re_activate(self, new_track, frame_id, new_id=False, update_feature=True):
        if self.use_kalman:
            self.mean, self.covariance = self.kalman_filter.update(self.mean, self.covariance, tlwh_to_xyah(new_track.tlwh))
        else:
            self.mean, self.covariance = None, None
            self._tlwh = np.asarray(new_track.tlwh, dtype=np.float64)
        if update_feature:
            self.update_features(new_track.curr_feat)
        self.tracklet_len = 0
        self.state = TrackState.Tracked
        self.is_activated = True
        self.frame_id = frame_id
        if new_id:
            self.track_id = self.next_id()
        if new_track.mask is not None:
            self.mask = new_track.mask

    def update(self, new_track, frame_id, update_feature=True):
        self.frame_id = frame_id
        self.tracklet_len += 1
        new_tlwh = new_track.tlwh
        if self.use_kalman:
            self.mean, self.covariance = self.kalman_filter.update(self.mean, self.covariance, tlwh_to_xyah(new_tlwh))
        else:
            self.mean, self.covariance = None, None
            self._tlwh = np.asarray(new_tlwh, dtype=np.float64)
        self.state = TrackState.Tracked
        self.is_activated = True
        self.score = new_track.score
        self.category = new_track.category
        if update_feature:
            self.update_features(new_track.curr_feat)
        if new_track.mask is not None:
            self.mask = new_track.mask
        if new_track.pose is not None:
            self.pose = new_track.pose

    @property
    def tlwh(self):
        if self.mean is None:
            return self._tlwh.copy()
        ret = self.mean[:4].copy()
        ret[2] *= ret[3]
        ret[:2] -= ret[2:] / 2
        return ret

    @property
    def tlbr(self):
        ret = self.tlwh.copy()
        ret[2:] += ret[:2]
        return ret

    def to_xyah(self):
        return tlwh_to_xyah(self.tlwh)

    def __repr__(self):
        return 'OT_{}_({}-{})'.format(self.track_id, self.start_frame, self.end_frame)


def joint_stracks(tlista, tlistb):
    exists, res = {}, []
    for t in tlista:
        exists[t.track_id] = 1
        res.append(t)
    for t in tlistb:
        if not exists.get(t.track_id, 0):
            exists[t.track_id] = 1
            res.append(t)
    return res


def sub_stracks(tlista, tlistb):
    stracks = {t.track_id: t for t in tlista}
    for t in tlistb:
        if stracks.get(t.track_id, 0):
            del stracks[t.track_id]
    return list(stracks.values())


def remove_duplicate_stracks(stracksa, stracksb, ioudist=0.15):
    pdist = iou_distance(stracksa, stracksb)
    dupa, dupb = [], []
    for p, q in zip(*np.where(pdist < ioudist)):
        timep = stracksa[p].frame_id - stracksa[p].start_frame
        timeq = stracksb[q].frame_id - stracksb[q].start_frame
        (dupb if timep > timeq else dupa).append(q if timep > timeq else p)
    return [t for i, t in enumerate(stracksa) if i not in dupa], [t for i, t in enumerate(stracksb) if i not in dupb]


class QueryFeatTube:
    """data/query_feat_tracklet.py."""

    def __init__(self, start_frame_id, track_id, query_feat):
        self.track_id = track_id
        self.start_frame_id = start_frame_id
        self.end_frame_id = start_frame_id
        self.len = 1
        self.qf_tube = [None for _ in range(self.start_frame_id - 1)] + [query_feat]

    def __repr__(self):
        return 'QFT_{}_({}_{})'.format(self.track_id, self.start_frame_id, self.end_frame_id)

    def update(self, query_feat, cur_frame_id):
        if self.end_frame_id < cur_frame_id:
            self.qf_tube.extend([None for _ in range(cur_frame_id - self.end_frame_id - 1)])
        self.qf_tube.append(query_feat)
        self.end_frame_id = cur_frame_id
        self.len += 1

    def complete_empty_postfix(self, last_frame_idx):
        if len(self.qf_tube) != last_frame_idx + 1:
            self.qf_tube.extend([None for _ in range(last_frame_idx + 1 - self.end_frame_id)])
        return self


# ======================================================================================
# core/association/matching.py
# ======================================================================================
def bbox_ious(boxes, query_boxes):
    """cython_bbox.bbox_overlaps (third-party, not vendored by the reference): IoU with the +1 pixel convention."""
    a, b = np.asarray(boxes, np.float64)[:, None, :], np.asarray(query_boxes, np.float64)[None, :, :]
    iw = np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0]) + 1
    ih = np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1]) + 1
    area = lambda t: (t[..., 2] - t[..., 0] + 1) * (t[..., 3] - t[..., 1] + 1)  # noqa: E731
    inter = np.where((iw > 0) & (ih > 0), iw * ih, 0.0)
    return np.where(inter > 0, inter / (area(a) + area(b) - inter), 0.0)


def ious(atlbrs, btlbrs):
    if len(atlbrs) == 0 or len(btlbrs) == 0:
        return np.zeros((len(atlbrs), len(btlbrs)), dtype=np.float64)
    return bbox_ious(np.ascontiguousarray(atlbrs, dtype=np.float64), np.ascontiguousarray(btlbrs, dtype=np.float64))


def iou_distance(atracks, btracks):
    if (len(atracks) > 0 and isinstance(atracks[0], np.ndarray)) or (len(btracks) > 0 and isinstance(btracks[0], np.ndarray)):
        atlbrs, btlbrs = atracks, btracks
    else:
        atlbrs, btlbrs = [t.tlbr for t in atracks], [t.tlbr for t in btracks]
    return 1 - ious(atlbrs, btlbrs)


def linear_assignment(cost_matrix, thresh):
    """matching.py:29-40 with ``lap.lapjv(extend_cost=True, cost_limit=thresh)`` solved by pvsg_lap_assign."""
    cost_matrix = np.asarray(cost_matrix)
    if cost_matrix.size == 0:
        return np.empty((0, 2), dtype=int), tuple(range(cost_matrix.shape[0])), tuple(range(cost_matrix.shape[1]))
    dev = torch.device('cuda', torch.cuda.current_device())
    x, y = ops.lap_assign(torch.as_tensor(cost_matrix, dtype=torch.float32).to(dev), float(thresh))
    x, y = x.cpu().numpy(), y.cpu().numpy()
    matches = np.asarray([[ix, mx] for ix, mx in enumerate(x) if mx >= 0])
    return matches, np.where(x < 0)[0], np.where(y < 0)[0]


def get_track_feat(tracks, feat_flag='curr'):
    """matching.py:170-191, position-major: [n, max positions, d] (the kernel's layout) instead of [n, d, positions]."""
    if feat_flag not in ('curr', 'smooth'):
        raise NotImplementedError
    feats = [(t.curr_feat if feat_flag == 'curr' else t.smooth_feat).squeeze(0) for t in tracks]
    feats = [f.reshape(f.shape[0], -1) for f in feats]
    width = max(f.shape[1] for f in feats)
    out = torch.zeros(len(feats), width, feats[0].shape[0], device=feats[0].device, dtype=torch.float32)
    for i, f in enumerate(feats):
        out[i, :f.shape[1]] = f.t()            # pure data movement
    return out


def reconsdot_distance(tracks, detections, tmp=100):
    """matching.py:194-238 -> (cost [ntrk, ndet] float64 numpy, None)."""
    cost_matrix = np.zeros((len(tracks), len(detections)), dtype=np.float64)
    if cost_matrix.size == 0:
        return cost_matrix, None
    dev = torch.device('cuda', torch.cuda.current_device())
    cost = ops.reconsdot(get_track_feat(tracks).to(dev), get_track_feat(detections).to(dev), tmp)
    return cost.cpu().numpy().astype(np.float64), None


def fuse_motion(kf, cost_matrix, tracks, detections, only_position=False, lambda_=0.98, gate=True):
    if cost_matrix.size == 0:
        return cost_matrix
    gating_threshold = chi2inv95[2 if only_position else 4]
    measurements = np.asarray([det.to_xyah() for det in detections])
    for row, track in enumerate(tracks):
        gating_distance = kf.gating_distance(track.mean, track.covariance, measurements, only_position, metric='maha')
        if gate:
            cost_matrix[row, gating_distance > gating_threshold] = np.inf
        cost_matrix[row] = lambda_ * cost_matrix[row] + (1 - lambda_) * gating_distance
    return cost_matrix


def category_gate(cost_matrix, tracks, detections):
    if cost_matrix.size == 0:
        return cost_matrix
    det_categories = np.array([d.category for d in detections])
    trk_categories = np.array([t.category for t in tracks])
    cost_matrix = cost_matrix + np.abs(det_categories[None, :] - trk_categories[:, None])
    return cost_matrix


def class_aware_distance(tracks, detections, query_feats):
    """multitracker.py:27-34."""
    dists, _ = reconsdot_distance(tracks, detections)
    for i, track in enumerate(tracks):
        for j, _det in enumerate(detections):
            if track.cls_id != query_feats[j]['cls_id'] % 1000:
                dists[i, j] = float('inf')
    return dists


# ======================================================================================
# multitracker.py / mask.py
# ======================================================================================
class AssociationTracker:
    """multitracker.py:36-205.  ``tracker_cfg``: the reference's config node (attribute access: ``.mots.*``, ``.common.*``)."""

    def __init__(self, tracker_cfg, app_model=None):
        self.tracker_cfg = tracker_cfg
        self.tracked_stracks, self.lost_stracks, self.removed_stracks = [], [], []
        self.query_feat_tubes = []      # always sorted by track_id
        self.frame_id = 0
        self.det_thresh = tracker_cfg.mots.conf_thres
        self.buffer_size = tracker_cfg.mots.track_buffer
        self.max_time_lost = self.buffer_size
        self.kalman_filter = KalmanFilter()
        self.app_model = app_model
        if not self.tracker_cfg.mots.asso_with_motion:
            self.tracker_cfg.mots.motion_lambda = 1
            self.tracker_cfg.mots.motion_gated = False

    def extract_emb(self, img, obs):
        raise NotImplementedError

    def prepare_obs(self, img, img0, obs, embs=None):
        raise NotImplementedError

    def _matched(self, track, det, query_feat, total_num_tubes_previous, activated, refind):
        self.query_feat_tubes[track.track_id - 1 - total_num_tubes_previous].update(query_feat, self.frame_id)
        if track.state == TrackState.Tracked:
            track.update(det, self.frame_id)
            activated.append(track)
        else:
            track.re_activate(det, self.frame_id, new_id=False)
            refind.append(track)

    def update(self, img, img0, obs, query_feats, total_num_tubes_previous, yembs=None):
        cfg = self.tracker_cfg.mots
        self.frame_id += 1
        activated_stracks, refind_stracks, lost_stracks, removed_stracks = [], [], [], []
        detections = self.prepare_obs(img, img0, obs, embs=None)
        unconfirmed = [t for t in self.tracked_stracks if not t.is_activated]
        tracked_stracks = [t for t in self.tracked_stracks if t.is_activated]

        # Step 2: first association, with the appearance embedding (class-aware)
        tracks = joint_stracks(tracked_stracks, self.lost_stracks)
        dists = class_aware_distance(tracks, detections, query_feats)
        if cfg.use_kalman:
            STrack.multi_predict(tracks)
            dists = fuse_motion(self.kalman_filter, dists, tracks, detections, lambda_=cfg.motion_lambda, gate=cfg.motion_gated)
        if obs.shape[1] == 6:
            dists = category_gate(dists, tracks, detections)
        matches, u_track, u_detection = linear_assignment(dists, thresh=0.9)
        for itracked, idet in matches:
            self._matched(tracks[itracked], detections[idet], query_feats[idet], total_num_tubes_previous, activated_stracks,
                          refind_stracks)

        if cfg.use_kalman:
            # Step 3: second association, with IoU
            tracks = [tracks[i] for i in u_track if tracks[i].state == TrackState.Tracked]
            detections = [detections[i] for i in u_detection]
            query_feats = [query_feats[i] for i in u_detection]
            dists = iou_distance(tracks, detections)
            matches, u_track, u_detection = linear_assignment(dists, thresh=0.5)
            for itracked, idet in matches:
                self._matched(tracks[itracked], detections[idet], query_feats[idet], total_num_tubes_previous,
                              activated_stracks, refind_stracks)
            # unconfirmed tracks (usually tracks with only one beginning frame)
            detections = [detections[i] for i in u_detection]
            query_feats = [query_feats[i] for i in u_detection]
            dists = iou_distance(unconfirmed, detections)
            matches, u_unconfirmed, u_detection = linear_assignment(dists, thresh=cfg.confirm_iou_thres)
            for itracked, idet in matches:
                unconfirmed[itracked].update(detections[idet], self.frame_id)
                activated_stracks.append(unconfirmed[itracked])
                self.query_feat_tubes[unconfirmed[itracked].track_id - 1 - total_num_tubes_previous].update(query_feats[idet],
                                                                                                             self.frame_id)
            for it in u_unconfirmed:
                unconfirmed[it].mark_removed()
                removed_stracks.append(unconfirmed[it])

        for it in u_track:
            track = tracks[it]
            if not track.state == TrackState.Lost:
                track.mark_lost()
                lost_stracks.append(track)

        # Step 4: new tracks
        for inew in u_detection:
            track = detections[inew]
            if track.score < self.det_thresh:
                continue
            track.activate(self.kalman_filter, self.frame_id)
            self.query_feat_tubes.append(QueryFeatTube(self.frame_id, track.track_id, query_feats[inew]))
            track.cls_id = query_feats[inew]['cls_id'] % 1000
            activated_stracks.append(track)

        # Step 5: state update
        for track in self.lost_stracks:
            if self.frame_id - track.end_frame > self.max_time_lost:
                track.mark_removed()
                removed_stracks.append(track)
        self.tracked_stracks = [t for t in self.tracked_stracks if t.state == TrackState.Tracked]
        self.tracked_stracks = joint_stracks(self.tracked_stracks, activated_stracks)
        self.tracked_stracks = joint_stracks(self.tracked_stracks, refind_stracks)
        self.lost_stracks = sub_stracks(self.lost_stracks, self.tracked_stracks)
        self.lost_stracks.extend(lost_stracks)
        self.lost_stracks = sub_stracks(self.lost_stracks, self.removed_stracks)
        self.removed_stracks.extend(removed_stracks)
        self.tracked_stracks, self.lost_stracks = remove_duplicate_stracks(self.tracked_stracks, self.lost_stracks,
                                                                           ioudist=cfg.dup_iou_thres)
        self.query_feat_tubes = sorted(self.query_feat_tubes, key=lambda q: q.track_id)
        output_stracks = [track for track in self.tracked_stracks if track.is_activated]
        return output_stracks, len(self.query_feat_tubes)

    def reset_all(self):
        self.tracked_stracks, self.lost_stracks, self.removed_stracks = [], [], []
        self.frame_id = 0


def coords2bbox(coords, extend=2):
    """utils/mask.py:18-37: centre +- ``extend`` x mean absolute deviation (at least 1) of the mask's pixel coordinates
    [(row, col)]; returned as (x0, y0, x1, y1)."""
    coords = np.asarray(coords, np.float32)
    center = coords.mean(0, dtype=np.float32)
    dis_r = max(np.float32(np.abs(coords[:, 0] - center[0]).mean(dtype=np.float32)), 1)
    dis_c = max(np.float32(np.abs(coords[:, 1] - center[1]).mean(dtype=np.float32)), 1)
    return (float(center[1] - dis_c * extend), float(center[0] - dis_r * extend),
            float(center[1] + dis_c * extend), float(center[0] + dis_r * extend))


def mask2box(masks):
    """utils/mask.py:69-78: masks [n, 1, h, w]; an empty mask gets the placeholder box (-1, -1, 10, 10)."""
    boxes = []
    for mask in masks:
        m = torch.nonzero(mask[0]).float().cpu().numpy()
        boxes.append(coords2bbox(m, extend=2) if m.size > 0 else (-1, -1, 10, 10))
    return np.asarray(boxes)


def remove_duplicated_box(boxes, iou_th=0.5):
    """utils/box.py:137-150: drop placeholder boxes, then every kept box (in order) discards all boxes overlapping it
    by more than ``iou_th`` (torchvision ``box_iou`` convention: no +1)."""
    b = np.asarray(boxes, np.float64).reshape(-1, 4)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = np.maximum(b[:, None, :2], b[None, :, :2])
    rb = np.minimum(b[:, None, 2:], b[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    inter = wh[..., 0] * wh[..., 1]
    with np.errstate(divide='ignore', invalid='ignore'):
        jac = (inter / (area[:, None] + area[None, :] - inter)).astype(np.float32) - np.eye(len(b), dtype=np.float32)
    keep = np.ones(len(b), bool)
    keep[(b[:, 0] == -1) & (b[:, 1] == -1) & (b[:, 2] == 10) & (b[:, 3] == 10)] = False
    for r in range(len(b)):
        if keep[r]:
            keep[jac[r] > iou_th] = False
    return np.where(keep)[0]


class MaskAssociationTracker(AssociationTracker):
    """mask.py:17-60: detections = instance masks; embedding = appearance features at the mask's pixels."""

    @torch.no_grad()
    def extract_emb(self, img, obs):
        dev = torch.device('cuda', torch.cuda.current_device())
        feat = self.app_model(img.unsqueeze(0).to(dev).float())                 # [1, d, h, w]
        _, d, h, w = feat.shape
        obs = torch.as_tensor(np.asarray(obs)).to(dev).float()
        ys = (torch.arange(h, device=dev) * obs.shape[1] // h)                  # F.interpolate(mode='nearest') = index select
        xs = (torch.arange(w, device=dev) * obs.shape[2] // w)
        obs = obs[:, ys][:, :, xs].unsqueeze(1)                                 # [n, 1, h, w]
        template_scale = int(np.prod(self.tracker_cfg.mots.feat_size))
        max_area = self.tracker_cfg.mots.max_mask_area
        feat_t = feat[0].permute(1, 2, 0).contiguous()                          # token-major [h, w, d]
        embs = []
        for ob in obs:
            scale = float(ob.sum())
            if scale <= 0:
                embs.append(torch.randn(d, template_scale))
                continue
            on = ob[0] > 0
            if scale > max_area:
                # masked features rescaled so that at most ~max_mask_area positions remain
                sf = float(np.sqrt(max_area / scale))
                oh, ow = int(np.floor(h * sf)), int(np.floor(w * sf))
                masked = torch.where(on[..., None], feat_t, torch.zeros((), device=dev))
                small = ops.bilinear_resize_scaled(masked[None], (oh, ow), 1.0 / sf)[0]      # [oh, ow, d]
                yy = torch.clamp((torch.arange(oh, device=dev).float() * (1.0 / sf)).floor().long(), max=h - 1)
                xx = torch.clamp((torch.arange(ow, device=dev).float() * (1.0 / sf)).floor().long(), max=w - 1)
                keep = on[yy][:, xx]
                emb = small[keep].t()
            else:
                emb = feat_t[on].t()
            embs.append(emb[None].cpu())                                        # [1, d, n_pix]
        return obs, embs

    def prepare_obs(self, img, img0, obs, embs=None):
        if obs.shape[0] == 0:
            return []
        masks, embs = self.extract_emb(img, obs)
        boxes = mask2box(masks)
        keep_idx = remove_duplicated_box(boxes, iou_th=0.7)
        return [STrack(tlbr_to_tlwh(boxes[k]), 1, embs[k], self.buffer_size, obs[k], ac=True) for k in keep_idx]


# ======================================================================================
# clip driver: data/single_video.py::LoadOutputsFromMask2Former + test_mots_from_mask2former.py::eval_seq
# ======================================================================================
def frame_observations(pan_mask, query_feat_dict, num_classes):
    """single_video.py:49-85: one binary mask per panoptic id of the frame (void = ``num_classes`` dropped, ids in the
    sorted order of ``np.unique``), with its query feature (the mean over the merged queries of a stuff segment) and
    class id (``id % INSTANCE_OFFSET``).  -> (obs int array [n,H,W] or empty, [dict(query_feat, cls_id)])."""
    from .mask2former import INSTANCE_OFFSET
    object_ids = [i for i in np.unique(pan_mask).tolist() if i != num_classes]
    if not object_ids:
        return np.array([]), []
    assert len(query_feat_dict) == len(object_ids), 'Masks and query feats should match!'
    masks, feats = [], []
    for oid in object_ids:
        masks.append((pan_mask == oid).astype(np.int64))
        qf = [np.asarray(x).squeeze() for x in query_feat_dict[oid]]
        feats.append(dict(query_feat=qf[0] if len(qf) == 1 else np.stack(qf).mean(axis=0), cls_id=oid % INSTANCE_OFFSET))
    return np.stack(masks), feats


def track_clip(outputs, frames, tracker_cfg, num_classes, app_model):
    """eval_seq (test_mots_from_mask2former.py:29-95) without the file / plotting side effects: per-frame IPS results
    (``pan_results``, ``query_feats``) -> MaskAssociationTracker.update -> (results, query_feat_tubes).

    ``frames``: the normalised frames the appearance network sees ([3,H,W] tensors, ``tracker_cfg.common.im_mean/std``
    already applied), one per output.  ``results`` rows: (frame id (1-based), tlwhs * down_factor, masks as dicts
    ``{size, counts (COCO RLE string), class_id}``, track ids) -- what ``write_mots_results`` writes to masks.txt."""
    from . import tubes
    BaseTrack.reset_count()
    tracker = MaskAssociationTracker(tracker_cfg, app_model)
    results = []
    frame_id = -1
    for frame_id, (out, img) in enumerate(zip(outputs, frames)):
        obs, query_feats = frame_observations(np.asarray(out['pan_results']), out['query_feats'], num_classes)
        if len(obs) == 0:
            results.append((frame_id + 1, [], [], []))
            continue
        targets, _ = tracker.update(img, None, obs, query_feats, 0)
        tlwhs, ids, masks = [], [], []
        for t in targets:
            m = np.asarray(t.mask).astype(np.uint8)
            masks.append(dict(size=list(m.shape), counts=tubes.rle_string(tubes.rle_counts(m)), class_id=t.cls_id))
            tlwhs.append(t.tlwh * tracker_cfg.common.down_factor)
            ids.append(t.track_id)
        results.append((frame_id + 1, tlwhs, masks, ids))
    tubes_out = [t.complete_empty_postfix(frame_id) for t in tracker.query_feat_tubes]
    return results, tubes_out


def mots_rows(results):
    """write_mots_results (utils/io.py:14-37): the lines of quantitive/masks.txt."""
    return [f'{fid} {tid} {rle["class_id"]} {rle["size"][0]} {rle["size"][1]} {rle["counts"]}'
            for fid, _, rles, tids in results for rle, tid in zip(rles, tids) if tid >= 0]
