from .ipb2dmapping import (compute_far_bound, compute_far_bound0406, compute_far_bound0606, find_aabb_box, kitti_dataload,
                           maicity_dataload, pack_train_rays)

# nof/dataset/__init__.py:3-6 -- train_kitti.py:50 looks the dataset class up by --datasettype
nof_dataset = {
    'kitti_dataload': kitti_dataload,
    'maicity_dataload': maicity_dataload,
}
