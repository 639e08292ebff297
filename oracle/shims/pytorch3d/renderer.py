class _Absent:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("pytorch3d.renderer stub: rasterisation is data preparation, not hot path")


class PerspectiveCameras(_Absent):
    pass


class MeshRasterizer(_Absent):
    pass


class RasterizationSettings(_Absent):
    pass
