"""Output containers of DD3DB200.forward.

When detectron2 / the reference's tridet package are importable (the deployment case: the reference's evaluators
drive this model) their own ``Instances`` / ``Boxes`` / ``Boxes3D`` are used, so downstream code sees the exact
types it expects (reference fields: fcos2d.py:331-335,263; fcos3d.py:398-399; consumer kitti_3d_evaluator.py:82-118).
Otherwise the minimal API-compatible containers below are used (same field names, indexing, ``cat``).
"""
import torch

try:  # pragma: no cover - not installable in the build container
    from detectron2.structures import Boxes, Instances  # type: ignore
    _HAVE_D2 = True
except Exception:  # noqa: BLE001
    _HAVE_D2 = False

if not _HAVE_D2:

    class Boxes:
        """detectron2.structures.Boxes subset: (N, 4) xyxy float tensor."""
        def __init__(self, tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
            if tensor.numel() == 0:
                tensor = tensor.reshape((-1, 4))
            assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
            self.tensor = tensor

        def clone(self):
            return Boxes(self.tensor.clone())

        def to(self, *args, **kwargs):
            return Boxes(self.tensor.to(*args, **kwargs))

        def area(self):
            b = self.tensor
            return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

        def nonempty(self, threshold=0.0):
            b = self.tensor
            return ((b[:, 2] - b[:, 0]) > threshold) & ((b[:, 3] - b[:, 1]) > threshold)

        def __getitem__(self, item):
            if isinstance(item, int):
                return Boxes(self.tensor[item].view(1, -1))
            return Boxes(self.tensor[item])

        def __len__(self):
            return self.tensor.shape[0]

        def __iter__(self):
            yield from self.tensor

        @classmethod
        def cat(cls, boxes_list):
            if len(boxes_list) == 0:
                return cls(torch.empty(0))
            return cls(torch.cat([b.tensor for b in boxes_list], dim=0))

        @property
        def device(self):
            return self.tensor.device

    class Instances:
        """detectron2.structures.Instances subset: equal-length named fields of one image."""
        def __init__(self, image_size, **kwargs):
            self._image_size = image_size
            self._fields = {}
            for k, v in kwargs.items():
                self.set(k, v)

        @property
        def image_size(self):
            return self._image_size

        def __setattr__(self, name, val):
            if name.startswith("_"):
                super().__setattr__(name, val)
            else:
                self.set(name, val)

        def __getattr__(self, name):
            if name == "_fields" or name not in self._fields:
                raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
            return self._fields[name]

        def set(self, name, value):
            if len(self._fields):
                assert len(self) == len(value), "Adding a field of length {} to a Instances of length {}".format(
                    len(value), len(self))
            self._fields[name] = value

        def has(self, name):
            return name in self._fields

        def get(self, name):
            return self._fields[name]

        def get_fields(self):
            return self._fields

        def to(self, *args, **kwargs):
            ret = Instances(self._image_size)
            for k, v in self._fields.items():
                ret.set(k, v.to(*args, **kwargs) if hasattr(v, "to") else v)
            return ret

        def __getitem__(self, item):
            if type(item) == int:
                if item >= len(self) or item < -len(self):
                    raise IndexError("Instances index out of range!")
                item = slice(item, None, len(self))
            ret = Instances(self._image_size)
            for k, v in self._fields.items():
                ret.set(k, v[item])
            return ret

        def __len__(self):
            for v in self._fields.values():
                return len(v)
            raise NotImplementedError("Empty Instances does not support __len__!")

        @staticmethod
        def cat(instance_lists):
            assert len(instance_lists) > 0
            if len(instance_lists) == 1:
                return instance_lists[0]
            ret = Instances(instance_lists[0].image_size)
            for k in instance_lists[0]._fields.keys():
                values = [i.get(k) for i in instance_lists]
                v0 = values[0]
                if isinstance(v0, torch.Tensor):
                    values = torch.cat(values, dim=0)
                elif hasattr(type(v0), "cat"):
                    values = type(v0).cat(values)
                else:
                    raise ValueError("Unsupported type {} for concatenation".format(type(v0)))
                ret.set(k, values)
            return ret


try:  # pragma: no cover
    from tridet.structures.boxes3d import Boxes3D  # type: ignore
except Exception:  # noqa: BLE001

    class Boxes3D:
        """tridet.structures.boxes3d.Boxes3D subset (boxes3d.py:157-289): vision-based 3-D boxes whose translation
        is derived lazily from the projected centre, depth and inverse intrinsics."""
        def __init__(self, quat, proj_ctr, depth, size, inv_intrinsics):
            self.quat = quat
            self.proj_ctr = proj_ctr
            self.depth = depth
            self.size = size
            self.inv_intrinsics = inv_intrinsics

        @property
        def tvec(self):  # boxes3d.py:169-173, geometry.py:86-112
            ones = torch.ones_like(self.proj_ctr[:, :1])
            ph = torch.cat([self.proj_ctr, ones], dim=1).unsqueeze(-1)
            ray = torch.matmul(self.inv_intrinsics, ph).squeeze(-1)
            return ray * self.depth

        def vectorize(self):  # boxes3d.py:142-144
            return torch.cat([self.quat, self.tvec, self.size], dim=1)

        @property
        def device(self):
            return self.quat.device

        def to(self, *args, **kwargs):
            return Boxes3D(self.quat.to(*args, **kwargs), self.proj_ctr.to(*args, **kwargs),
                           self.depth.to(*args, **kwargs), self.size.to(*args, **kwargs),
                           self.inv_intrinsics.to(*args, **kwargs))

        def clone(self):
            return Boxes3D(self.quat.clone(), self.proj_ctr.clone(), self.depth.clone(), self.size.clone(),
                           self.inv_intrinsics.clone())

        def __getitem__(self, item):
            if isinstance(item, int):
                return Boxes3D(self.quat[item].view(1, -1), self.proj_ctr[item].view(1, -1),
                               self.depth[item].view(1, -1), self.size[item].view(1, -1),
                               self.inv_intrinsics[item].view(1, 3, 3))
            return Boxes3D(self.quat[item], self.proj_ctr[item], self.depth[item], self.size[item],
                           self.inv_intrinsics[item])

        def __len__(self):
            return self.quat.shape[0]

        @classmethod
        def cat(cls, boxes_list, dim=0):
            if len(boxes_list) == 0:
                return cls(torch.empty(0), torch.empty(0), torch.empty(0), torch.empty(0), torch.empty(0))
            return cls(*[torch.cat([getattr(b, f) for b in boxes_list], dim=dim)
                         for f in ("quat", "proj_ctr", "depth", "size", "inv_intrinsics")])


try:  # pragma: no cover
    from tridet.structures.boxes3d import GenericBoxes3D  # type: ignore
except Exception:  # noqa: BLE001

    class GenericBoxes3D:
        """tridet.structures.boxes3d.GenericBoxes3D subset (boxes3d.py:19-144): (quat wxyz, tvec, size WLH) rows; the
        type of ``pred_boxes3d_global`` (postprocessing.py:50-51)."""
        def __init__(self, quat, tvec, size):
            self.quat, self._tvec, self.size = quat, tvec, size

        @property
        def tvec(self):
            return self._tvec

        def vectorize(self):
            return torch.cat([self.quat, self.tvec, self.size], dim=1)

        @property
        def device(self):
            return self.quat.device

        def to(self, *args, **kwargs):
            return GenericBoxes3D(self.quat.to(*args, **kwargs), self.tvec.to(*args, **kwargs),
                                  self.size.to(*args, **kwargs))

        def __getitem__(self, item):
            if isinstance(item, int):
                return GenericBoxes3D(self.quat[item].view(1, -1), self.tvec[item].view(1, -1), self.size[item].view(1, -1))
            return GenericBoxes3D(self.quat[item], self.tvec[item], self.size[item])

        def __len__(self):
            return self.quat.shape[0]

        @classmethod
        def cat(cls, boxes_list, dim=0):
            return cls(*[torch.cat([getattr(b, f) for b in boxes_list], dim=dim) for f in ("quat", "tvec", "size")])


def matrix_to_quaternion_wxyz(R):
    """3x3 rotation (tensor) -> [w, x, y, z] (host-side helper for 4x4 pose inputs)."""
    m = torch.as_tensor(R, dtype=torch.float64)
    tr = float(m[0, 0] + m[1, 1] + m[2, 2])
    if tr > 0:
        s = (tr + 1.0)**0.5 * 2
        q = [0.25 * s, float(m[2, 1] - m[1, 2]) / s, float(m[0, 2] - m[2, 0]) / s, float(m[1, 0] - m[0, 1]) / s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = float(1.0 + m[0, 0] - m[1, 1] - m[2, 2])**0.5 * 2
        q = [float(m[2, 1] - m[1, 2]) / s, 0.25 * s, float(m[0, 1] + m[1, 0]) / s, float(m[0, 2] + m[2, 0]) / s]
    elif m[1, 1] > m[2, 2]:
        s = float(1.0 + m[1, 1] - m[0, 0] - m[2, 2])**0.5 * 2
        q = [float(m[0, 2] - m[2, 0]) / s, float(m[0, 1] + m[1, 0]) / s, 0.25 * s, float(m[1, 2] + m[2, 1]) / s]
    else:
        s = float(1.0 + m[2, 2] - m[0, 0] - m[1, 1])**0.5 * 2
        q = [float(m[1, 0] - m[0, 1]) / s, float(m[0, 2] + m[2, 0]) / s, float(m[1, 2] + m[2, 1]) / s, 0.25 * s]
    return q
