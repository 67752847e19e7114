"""The data format behind the evaluation loop: posed target images of a dataset directory, in the layout
``PosedImage`` of the reference produces (data/datasets.py:377-713), ready for ``trainer.Trainer.test``.

``PosedViews(root_directory, stage)`` reads ``views/transforms_<stage>.json`` (NeRF-synthetic style:
``camera_angle_x`` or explicit ``intrinsics``, optional ``bit_depth``, per frame ``file_path``,
``transform_matrix`` and optionally ``exposure_time`` / ``gain``), the image files (OpenCV, unchanged: grey,
BGR or BGRA, 8 / 16-bit or float), ``renderer_params.npz`` (synthetic renders: intermediate colour space and
log epsilon) and the Bayer pattern of ``camera_calibration.npz``, and applies the reference's image and pose
transformations (:549-713) as elementwise tensor operations on `device`:

* BGRA renders are composited over a white background when ``alpha_over_white_bg`` (straight alpha in the
  display colour space, premultiplied in the linear one), else the alpha channel is dropped;
* a sensor behind a Bayer filter keeps colour, as (3, H, W) RGB; a monochrome sensor with three-channel images
  converts to grey with OpenCV's weights (0.299 R + 0.587 G + 0.114 B);
* quantised images map level x to (x + 0.5) / levels; float renders get the log epsilon added;
* camera poses go from the OpenGL camera frame (x right, y up, z backwards) to the common one (x right, y down,
  z forwards); sample ids are the file names as 16 unicode code points.

Iterating yields the per-view dicts ``Trainer.test`` / ``EventRenderer.evaluation_step`` take; `test_arguments()`
returns the rest of what that loop needs (inverse intrinsics, the pixel value range)."""

import glob
import json
import math
import os

import numpy as np
import torch

STAGES = ("train", "val", "test")
SAMPLE_ID_LEN = 16                                      # data/datasets.py:379
T_COPENGL_CCOMMON = np.array([[1, 0, 0], [0, -1, 0], [0, 0, -1]])       # :381-383
GREY_WEIGHTS_BGR = (0.114, 0.587, 0.299)                # cv2.COLOR_BGR2GRAY


def _views_folder(root_directory):
    for path in (os.path.join(root_directory, "views"), os.path.join(root_directory, "..", "views")):    # :435-445
        if os.path.isdir(path):
            return path
    raise FileNotFoundError(f"no views folder in or above {root_directory}")


class PosedViews:
    def __init__(self, root_directory, stage, permutation_seed=None, alpha_over_white_bg=False, device="cpu"):
        if stage not in STAGES:
            raise ValueError(f"stage must be one of {STAGES}")
        import cv2
        folder = _views_folder(root_directory)
        with open(os.path.join(folder, f"transforms_{stage}.json")) as fh:
            transforms = json.load(fh)
        params_path = os.path.join(root_directory, "renderer_params.npz")
        renderer = np.load(params_path) if os.path.isfile(params_path) else None
        calib = np.load(os.path.join(root_directory, "camera_calibration.npz"))
        frames = transforms["frames"]

        ids, images, poses = [], [], []
        for frame in frames:                                                                        # :487-521
            name = os.path.basename(frame["file_path"]).ljust(SAMPLE_ID_LEN)
            ids.append([ord(c) for c in name])
            path = glob.glob(os.path.join(folder, frame["file_path"] + ".*"))[0]
            images.append(cv2.imread(path, cv2.IMREAD_UNCHANGED))
            poses.append(np.array(frame["transform_matrix"]))
        images = np.stack(images, axis=0)
        poses = np.stack(poses, axis=0)
        self.sample_id = torch.tensor(ids)
        if "camera_angle_x" in transforms:                                                          # :531-541
            h, w = images.shape[1:3]
            focal = (w / 2) / math.tan(transforms["camera_angle_x"] / 2)
            intrinsics = np.array([[focal, 0, w / 2 - 0.5], [0, focal, h / 2 - 0.5], [0, 0, 1]])
        else:
            intrinsics = np.array(transforms["intrinsics"])

        self.img = self._transform_images(images, transforms, renderer, str(calib["bayer_pattern"]),
                                          alpha_over_white_bg, torch.device(device))
        dtype = torch.get_default_dtype()
        orientation = poses[:, :3, :3] @ T_COPENGL_CCOMMON                                          # :697-699
        self.T_wc_position = torch.tensor(poses[:, :3, 3], dtype=dtype)
        self.T_wc_orientation = torch.tensor(orientation, dtype=dtype)
        self.intrinsics = torch.tensor(intrinsics, dtype=dtype)
        self.exposure_time = self.gain = None
        if frames and "exposure_time" in frames[0]:
            self.exposure_time = torch.tensor([f["exposure_time"] for f in frames])
        if frames and "gain" in frames[0]:
            self.gain = torch.tensor([f["gain"] for f in frames], dtype=dtype)
        if permutation_seed is not None:                                                            # :425-433
            generator = torch.Generator()
            generator.manual_seed(permutation_seed)
            perm = torch.randperm(len(self.img), generator=generator)
            for key in self._per_view_keys():
                value = getattr(self, key)
                setattr(self, key, value[perm.to(value.device)])

    def _transform_images(self, images, transforms, renderer, bayer_pattern, alpha_over_white_bg, device):
        """:549-682 as tensor operations.  images (N, H, W [, 3 / 4]) as read from disk."""
        quantised = np.issubdtype(images.dtype, np.unsignedinteger)
        synthetic = renderer is not None
        channels = 1 if images.ndim == 3 else images.shape[3]
        if not (quantised or np.issubdtype(images.dtype, np.floating)) or (images < 0).any():
            raise ValueError("images must hold unsigned integers or non-negative floats")
        if channels not in (1, 3, 4) or (channels == 4 and not synthetic) or not (synthetic or quantised):
            raise ValueError("unsupported image format (grey / BGR / BGRA; an alpha channel or float pixels only "
                             "for synthetic renders)")
        space = str(renderer["interm_color_space"]) if synthetic else None
        if synthetic and space != ("display" if quantised else "linear"):
            raise ValueError("quantised renders must be in the display colour space, float renders in the linear one")
        levels = None
        if quantised:
            levels = 2 ** transforms["bit_depth"] if "bit_depth" in transforms else np.iinfo(images.dtype).max + 1

        # float64 where upstream's numpy arithmetic is float64 (the alpha compositing of quantised images)
        img = torch.from_numpy(images.astype(np.float64 if quantised else images.dtype)).to(device)
        if alpha_over_white_bg:                                                                     # :611-629
            if space == "display":
                alpha = (img[..., 3] / (levels - 1)).unsqueeze(-1)
                img = alpha * img[..., :3] + (1 - alpha) * (levels - 1)
            elif space == "linear":
                img = img[..., :3] + (1 - img[..., 3].unsqueeze(-1))
        elif channels == 4:
            img = img[..., :3]
        img = img.to(torch.float32)                                                                 # :632
        if bayer_pattern != "":                                                                     # :636-641
            img = img.flip(-1).permute(0, 3, 1, 2)                  # BGR -> RGB, channels first
        elif channels == 3:                                                                         # :644-648
            b, g, r = GREY_WEIGHTS_BGR
            img = img[..., 0] * b + img[..., 1] * g + img[..., 2] * r
        if quantised:                                                                               # :660-666
            self.min_normalized_pixel_value = 0.5 / levels
            img = img / levels + self.min_normalized_pixel_value
            self.max_normalized_pixel_value = 1 - self.min_normalized_pixel_value
        else:                                                                                       # :669-672
            self.min_normalized_pixel_value = float(renderer["log_eps"])
            img = img + self.min_normalized_pixel_value
            self.max_normalized_pixel_value = float(img.max())
        return img.to(torch.get_default_dtype()).contiguous()

    def _per_view_keys(self):
        keys = ["sample_id", "img", "T_wc_position", "T_wc_orientation"]
        return keys + [k for k in ("exposure_time", "gain") if getattr(self, k) is not None]

    def __len__(self):
        return len(self.img)

    def __getitem__(self, index):
        """The reference's per-view dict (:704-709)."""
        return {key: getattr(self, key)[index] for key in self._per_view_keys()}

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def test_arguments(self, device=None):
        """What `trainer.Trainer.test(model, views, **arguments)` needs beside the views: the inverse
        intrinsics (models/deblur_e_nerf.py:108-119,146-152) and the pixel value range of the stage."""
        device = self.img.device if device is None else device
        return {"intrinsics_inv": torch.linalg.inv(self.intrinsics).to(device),
                "min_normalized_pixel_value": self.min_normalized_pixel_value,
                "max_normalized_pixel_value": self.max_normalized_pixel_value}


def sample_id_strings(sample_id):
    """(B, 16) unicode code points -> file names without the padding (models/deblur_e_nerf.py:1311-1319)."""
    return ["".join(map(chr, row.tolist())).rstrip() for row in sample_id]


def save_predictions(pred, sample_id, folder, min_normalized_pixel_value, max_normalized_pixel_value,
                     bit_depth=8):
    """`eval_save_pred_intensity_img` (models/deblur_e_nerf.py:1008-1054): the corrected predictions (B, C, H, W)
    — C = 1 grey, C = 3 RGB of a sensor behind a Bayer filter — normalised to the stage's pixel range, clipped,
    quantised to `bit_depth` bits on the device and written as `<folder>/<sample id>.png` (OpenCV: BGR order).
    Returns the file paths."""
    import cv2
    if bit_depth not in (8, 16):
        raise NotImplementedError("prediction bit depth must be 8 or 16")
    names = sample_id_strings(sample_id) if torch.is_tensor(sample_id) else list(sample_id)
    if len(names) != len(pred):
        raise ValueError("one sample id per predicted image")
    levels = 2 ** bit_depth - 1
    span = max_normalized_pixel_value - min_normalized_pixel_value
    q = (levels * ((pred - min_normalized_pixel_value) / span).clamp(min=0, max=1)).round()
    q = q.permute(0, 2, 3, 1).cpu().numpy().astype(np.uint8 if bit_depth == 8 else np.uint16)
    if q.shape[-1] == 3:
        q = q[..., ::-1]                                                                            # RGB -> BGR
    os.makedirs(folder, exist_ok=True)
    paths = []
    for name, img in zip(names, q):
        paths.append(os.path.join(folder, name + ".png"))
        cv2.imwrite(paths[-1], np.ascontiguousarray(img))
    return paths
