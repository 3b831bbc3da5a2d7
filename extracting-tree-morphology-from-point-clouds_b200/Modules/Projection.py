"""Drop-in for the reference's ``Modules/Projection.py`` (variant B of the kernel: perpendicularity
tolerance 1e-3, epsilon guards on ``||rejection||`` and ``||axis||``, ``move_points_to_mantle`` flag).

    closest_cylinder_cuda_batch          reference :19-115
    generate_offset_cloud_cuda_batched   reference :117-144
    fit_circle_2d / get_point_cloud_stem_base_center / get_qsm_stem_base_center   reference :149-258
    project_clouds                       reference :264-444

``QSMFittingDepthFirst.cylinder_proximity_based_segmentation`` (reference :1079-1081) imports
``closest_cylinder_cuda_batch`` from this module and passes pre-built torch tensors; that works here too.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

from .. import api, dropin
from .Features import add_features
from .Utils import get_device, load_cloud

VARIANT = api.VARIANT_B

# internal column name -> accepted CSV spellings, most preferred first (reference :287-296)
QSM_COLUMN_ALIASES = {
    "startX": ("startX", "start.x", "start_x"),
    "startY": ("startY", "start.y", "start_y"),
    "startZ": ("startZ", "start.z", "start_z"),
    "endX": ("endX", "end.x", "end_x"),
    "endY": ("endY", "end.y", "end_y"),
    "endZ": ("endZ", "end.z", "end_z"),
    "radius": ("radius", "Radius"),
    "ID": ("ID", "extension", "id"),
}


def closest_cylinder_cuda_batch(points, start, radius, axis_length, axis_unit, IDs, device, move_points_to_mantle=True):
    return dropin.closest_cylinder(points, start, radius, axis_length, axis_unit, IDs, device, VARIANT,
                                   move_points_to_mantle=move_points_to_mantle)


def generate_offset_cloud_cuda_batched(cloud, cylinders, device, masterBar=None, batch_size=1024):
    return dropin.offset_cloud(cloud, cylinders, device, VARIANT, masterBar=masterBar, batch_size=batch_size)


# ---- stem-base alignment helpers ------------------------------------------------------------------

def fit_circle_2d(points_2d):
    """Algebraic least-squares circle: returns (centre (2,), radius), NaNs when it cannot be fitted."""
    nan2 = np.array([np.nan, np.nan])
    if points_2d.shape[0] < 3:
        return nan2, np.nan
    x, y = points_2d[:, 0], points_2d[:, 1]
    design = np.c_[2 * x, 2 * y, np.ones_like(x)]
    try:
        (cx, cy, c), *_ = np.linalg.lstsq(design, x * x + y * y, rcond=None)
    except np.linalg.LinAlgError:
        return nan2, np.nan
    r2 = c + cx * cx + cy * cy
    if r2 < 0:
        return nan2, np.nan
    return np.array([cx, cy]), np.sqrt(r2)


def get_point_cloud_stem_base_center(cloud_xyz, slice_height_from_min_z=0.10, num_ransac_fits=5, ransac_subset_ratio=0.7):
    """[x, y, z_min] of the stem base: mean centre of circle fits to random subsets of the lowest slice."""
    if cloud_xyz.shape[0] < 10:
        print("[WARNING] PC stem base: Not enough points in cloud.")
        return None
    z_min = np.min(cloud_xyz[:, 2])
    z = cloud_xyz[:, 2]
    base = cloud_xyz[(z >= z_min) & (z < z_min + slice_height_from_min_z)]
    if base.shape[0] < 10:
        print(f"[WARNING] PC stem base: Not enough points ({base.shape[0]}) in slice "
              f"[{z_min:.2f}-{z_min + slice_height_from_min_z:.2f}]. Trying wider slice up to 0.5m.")
        base = cloud_xyz[z < z_min + 0.5]
        if base.shape[0] < 10:
            cx, cy = np.mean(cloud_xyz[:, :2], axis=0)
            print(f"[WARNING] PC stem base: Fallback to full cloud centroid XY [{cx:.2f}, {cy:.2f}] at min_Z {z_min:.2f}.")
            return np.array([cx, cy, z_min])
    xy = base[:, :2]
    n = xy.shape[0]
    take = min(n, max(3, int(n * ransac_subset_ratio)))
    centres = []
    if n >= 3:
        for _ in range(num_ransac_fits):
            pick = np.random.choice(n, size=take, replace=False)
            centre, _ = fit_circle_2d(xy[pick])
            if not np.any(np.isnan(centre)):
                centres.append(centre)
    centre = np.mean(np.array(centres), axis=0) if centres else np.array([np.nan, np.nan])
    if np.any(np.isnan(centre)):
        centre, _ = fit_circle_2d(xy)
        if np.any(np.isnan(centre)):
            print("[WARNING] PC stem base: All circle fits failed. Using mean XY of slice.")
            centre = np.mean(xy, axis=0)
            if np.any(np.isnan(centre)):
                print("[ERROR] PC stem base: Cannot determine XY center. Critical error.")
                return None
    return np.array([centre[0], centre[1], z_min])


def get_qsm_stem_base_center(qsm_df):
    """Start point [x, y, z] of the lowest stem cylinder (BranchOrder 0 if that column exists)."""
    needed = ["startZ", "startX", "startY"]
    if "BranchOrder" in qsm_df.columns:
        needed.append("BranchOrder")
    if qsm_df.empty or any(c not in qsm_df.columns for c in needed):
        print("[WARNING] QSM lowest stem: Missing required columns or empty DataFrame.")
        return None
    work = qsm_df.copy()
    try:
        for c in needed:
            work[c] = pd.to_numeric(work[c], errors="coerce")
        work = work.dropna(subset=needed)
    except Exception as exc:
        print(f"Error converting QSM columns: {exc}")
        return None
    if work.empty:
        print("[WARNING] QSM lowest stem: No valid QSM data after NaN drop.")
        return None
    candidates = work
    if "BranchOrder" in work.columns:
        stem = work[work["BranchOrder"] == 0]
        if stem.empty:
            print("[WARNING] QSM lowest stem: No BranchOrder 0 cylinders. Using all cylinders to find lowest Z.")
        else:
            candidates = stem
            print(f"[INFO] QSM lowest stem: Found {len(stem)} cylinders with BranchOrder 0.")
    else:
        print("[INFO] QSM lowest stem: BranchOrder not available. Using all cylinders to find lowest Z.")
    row = candidates.loc[candidates["startZ"].idxmin()]
    base = np.array([row["startX"], row["startY"], row["startZ"]], dtype=float)
    if np.any(np.isnan(base)):
        print("[WARNING] QSM lowest stem: Calculated base coordinates are NaN.")
        return None
    return base


# ---- driver ---------------------------------------------------------------------------------------

def _stem(path):
    return os.path.splitext(os.path.basename(path))[0]


def _match_qsm(cloud_stem, qsm_stems):
    """QSM whose base name starts with the cloud's base name, shortest suffix first (reference :299-314)."""
    best = None
    for stem, path in qsm_stems:
        if stem.startswith(cloud_stem) and (best is None or len(stem) < len(best[0])):
            best = (stem, path)
    return None if best is None else best[1]


def _standardise_columns(raw, qsm_path):
    """Map the CSV's spellings onto the internal column names; ``None`` if an essential one is missing."""
    table = pd.DataFrame()
    have = raw.columns.tolist()
    for name, spellings in QSM_COLUMN_ALIASES.items():
        hit = next((s for s in spellings if s in have), None)
        if hit is None:
            print(f"[⚠️ WARNING] QSM {qsm_path}: Could not find data for essential field '{name}'. "
                  f"Tried candidates: {list(spellings)}. Available CSV columns: {have}. Skipping file.")
            return None
        table[name] = raw[hit]
    return table


def _align_to_cloud(table, cloud_xyz, cloud_path):
    print(f"[INFO] Aligning QSM stem base to cloud stem base for: {os.path.basename(cloud_path)}")
    pc_ref = get_point_cloud_stem_base_center(cloud_xyz, slice_height_from_min_z=0.10)
    qsm_ref = get_qsm_stem_base_center(table.copy())
    if pc_ref is None or qsm_ref is None:
        print("[WARNING] Could not determine both PC and QSM stem base references. Skipping alignment for this file.")
        return table
    shift = qsm_ref - pc_ref
    print(f"  PC stem base ref (local): {pc_ref}")
    print(f"  QSM stem base ref (global): {qsm_ref}")
    print(f"  Calculated Translation Vector to SUBTRACT from QSM: {shift}")
    moved = table.copy()
    for k, axis in enumerate("XYZ"):
        for end in ("start", "end"):
            moved[f"{end}{axis}"] = pd.to_numeric(moved[f"{end}{axis}"], errors="coerce") - shift[k]
    return moved


def project_clouds(cloudList, cylinderList, labelDir, batch_size=1024, use_features=False, denoised=False,
                   align_qsm_to_cloud=False):
    device = get_device()
    suffix = "_labeled_pred_denoised_projected.npy" if denoised else "_labeled_pred_projected.npy"
    qsm_stems = [(_stem(p), p) for p in cylinderList]
    print("\nMatching and Labeling clouds...")
    done = 0
    for cloud_path in cloudList:
        cloud_stem = _stem(cloud_path)
        qsm_path = _match_qsm(cloud_stem, qsm_stems)
        if qsm_path is None:
            print(f"[⚠️ WARNING] No matching QSM found for cloud: {cloud_path} (Basename: {cloud_stem})")
            continue
        print(f"[INFO] Matching Cloud: {os.path.basename(cloud_path)}  ->  QSM: {os.path.basename(qsm_path)}")
        cloud = load_cloud(cloud_path)
        if cloud is None or cloud.shape[0] == 0:
            print(f"[⚠️ WARNING] Cloud {cloud_path} is empty or failed to load. Skipping.")
            continue
        if cloud.ndim != 2 or cloud.shape[1] < 3:
            print(f"[⚠️ WARNING] Cloud {cloud_path} does not have expected shape (N, >=3). Actual shape: {cloud.shape}. Skipping.")
            continue
        try:
            raw = pd.read_csv(qsm_path, header=0)
            raw.columns = raw.columns.str.strip().str.replace('"', "")
        except pd.errors.EmptyDataError:
            print(f"[⚠️ ERROR] QSM file {qsm_path} is empty or has no columns. Skipping projection for {cloud_path}.")
            continue
        except FileNotFoundError:
            print(f"[⚠️ ERROR] QSM file {qsm_path} not found. Skipping for {cloud_path}.")
            continue
        except Exception as exc:
            print(f"[⚠️ ERROR] Failed to read QSM {qsm_path}: {exc}. Skipping for {cloud_path}.")
            continue
        if raw.empty:
            print(f"[⚠️ WARNING] QSM {qsm_path} loaded but is empty. Skipping projection for {cloud_path}.")
            continue
        table = _standardise_columns(raw, qsm_path)
        if table is None:
            continue
        try:
            table["ID"] = table["ID"].astype(int)
        except ValueError as exc:
            print(f"[⚠️ WARNING] QSM {qsm_path}: Could not convert 'ID' column to integer ({exc}). Skipping file.")
            continue
        hollow = table.columns[table.isnull().all()].tolist()
        if hollow:
            print(f"[⚠️ WARNING] QSM {qsm_path}: After mapping, columns {hollow} are entirely NaN/empty. Skipping.")
            continue
        if align_qsm_to_cloud:
            table = _align_to_cloud(table, cloud[:, :3], cloud_path)

        os.makedirs(labelDir, exist_ok=True)
        target = os.path.join(labelDir, cloud_stem + suffix)
        if use_features:
            projected = generate_offset_cloud_cuda_batched(cloud, table, device, batch_size=batch_size)
            projected = add_features(projected, use_densities=False, use_curvatures=False, use_distances=False,
                                     use_verticalities=False)
            np.save(target, projected)
        else:       # the (N,11) rows (four columns of ones appended) are written once, straight into the .npy file
            dropin.offset_cloud_to_npy(target, cloud, table, device, VARIANT)
        done += 1
    print(f"\n✅ Finished labeling and saving! {done} cloud(s) processed.")
