"""Host-side mirror of tuch/train/fits_dict.py: FitsDict (:29-119) -- the per-image store of the best
SMPL fits.  The store itself stays what it is in the reference (one [N,82] CPU tensor per dataset,
saved as `<dataset>_fits.npy`); the per-batch pose transforms, which the reference runs on the host
with a cv2.Rodrigues call per sample, are one kernel (tuch_fits_pose_transform).

Where the reference reads options.checkpoint_dir / config.STATIC_FITS_DIR /
constants.SMPL_POSE_FLIP_PERM from its un-shipped tree, the same values can be handed in directly.
"""
import os

import numpy as np
import torch

from .. import ops

# public SPIN constant (data.essentials.constants.SMPL_POSE_FLIP_PERM): left/right joint swap, per axis
_FLIP_JOINTS = [0, 2, 1, 3, 5, 4, 6, 8, 7, 9, 11, 10, 12, 14, 13, 15, 17, 16, 19, 18, 21, 20, 23, 22]
SMPL_POSE_FLIP_PERM = [3 * j + k for j in _FLIP_JOINTS for k in range(3)]


class FitsDict():
    """ Dictionary keeping track of the best fit per image in the training set """

    def __init__(self, options=None, train_dataset=None, device=None, checkpoint_dir=None, static_fits_dir=None,
                 dataset_sizes=None, flip_perm=None):
        self.options = options
        self.train_dataset = train_dataset
        self.device = torch.device(device if device is not None else 'cuda')
        if self.device.type != 'cuda':
            raise ops.TuchError('FitsDict needs a CUDA device for its pose transforms: tuch_b200 has no CPU fallback')
        self.checkpoint_dir = checkpoint_dir if checkpoint_dir is not None else getattr(options, 'checkpoint_dir', None)
        self.flipped_parts = torch.tensor(SMPL_POSE_FLIP_PERM if flip_perm is None else flip_perm, dtype=torch.int64)
        self._perm_dev = self.flipped_parts.to(self.device, torch.int32)
        if dataset_sizes is None:
            dataset_sizes = {name: len(train_dataset.datasets[idx]) for name, idx in train_dataset.dataset_dict.items()}
        self.fits_dict = {}
        for ds_name, n in dataset_sizes.items():
            loaded = None
            for folder in (self.checkpoint_dir, static_fits_dir):
                f = os.path.join(folder, ds_name + '_fits.npy') if folder else None
                if f and os.path.isfile(f):
                    loaded = torch.from_numpy(np.load(f))
                    break
            if loaded is None:
                print('Warning no statis fits exists. Mean pose created.')
                loaded = torch.zeros((n, 82))
            self.fits_dict[ds_name] = loaded

    def save(self):
        """ Save dictionary state to disk """
        for ds_name, fits in self.fits_dict.items():
            np.save(os.path.join(self.checkpoint_dir, ds_name + '_fits.npy'), fits.cpu().numpy())

    def __getitem__(self, x):
        """ Retrieve dictionary entries: (pose[B,72], betas[B,10]) with rotation and flipping applied """
        dataset_name, ind, rot, is_flipped = x
        ind = torch.as_tensor(ind, dtype=torch.long).cpu()
        params = torch.empty(len(dataset_name), 82)
        for ds, pos in self._rows_by_dataset(dataset_name).items():       # one gather per dataset, not per sample
            params[pos] = self.fits_dict[ds][ind[pos]].float()
        params = params.to(self.device)
        pose = ops.fits_pose_transform(params[:, :72], rot, is_flipped, self._perm_dev, flip_first=False)
        return pose, params[:, 72:].clone()

    def __setitem__(self, x, val):
        """ Update dictionary entries """
        dataset_name, ind, rot, is_flipped, update = x
        pose, betas = val
        # undo flipping and rotation (fits_dict.py:83)
        pose = ops.fits_pose_transform(pose.to(self.device), -rot.to(self.device, torch.float32), is_flipped,
                                       self._perm_dev, flip_first=True)
        params = torch.cat((pose, betas.to(self.device)), dim=-1).cpu()
        ind = torch.as_tensor(ind, dtype=torch.long).cpu()
        update = torch.as_tensor(update).cpu().bool()
        for ds, pos in self._rows_by_dataset(dataset_name).items():
            sel = pos[update[pos]]
            if len(sel):
                self.fits_dict[ds][ind[sel]] = params[sel].to(self.fits_dict[ds].dtype)

    @staticmethod
    def _rows_by_dataset(dataset_name):
        rows = {}
        for n, ds in enumerate(dataset_name):
            rows.setdefault(ds, []).append(n)
        return {ds: torch.tensor(r, dtype=torch.long) for ds, r in rows.items()}

    def flip_pose(self, pose, is_flipped):
        """flip SMPL pose parameters"""
        return ops.fits_pose_transform(pose.to(self.device), None, is_flipped, self._perm_dev)

    def rotate_pose(self, pose, rot):
        """Rotate SMPL pose parameters by rot degrees"""
        return ops.fits_pose_transform(pose.to(self.device), rot, None, None)
