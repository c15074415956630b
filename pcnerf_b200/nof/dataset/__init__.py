from .ipb2dmapping import compute_far_bound, compute_far_bound0406, compute_far_bound0606, find_aabb_box, pack_train_rays
