#!/usr/bin/env python
"""Extract the reference's rot6d normalisation statistics into dposer_b200/data (build container only):
    python tests/golden/make_data_rot6d.py
Source: data/AMASS/amass_processed/version1/train/rot6d_normalize{1,2}.pt (lib/dataset/AMASS.py:193-196)."""
import os

import numpy as np
import torch

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    d = os.path.join(REF, 'data/AMASS/amass_processed/version1/train')
    p1, p2 = torch.load(os.path.join(d, 'rot6d_normalize1.pt')), torch.load(os.path.join(d, 'rot6d_normalize2.pt'))
    np.savez(os.path.join(ROOT, 'dposer_b200/data/amass_stats_rot6d.npz'),
             mean_poses=p2['mean_poses'].numpy(), std_poses=p2['std_poses'].numpy(),
             min_poses=p1['min_poses'].numpy(), max_poses=p1['max_poses'].numpy())
    print('wrote amass_stats_rot6d.npz')


if __name__ == '__main__':
    main()
