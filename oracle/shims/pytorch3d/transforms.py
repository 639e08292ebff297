def euler_angles_to_matrix(*a, **k):
    raise RuntimeError("pytorch3d.transforms stub")
