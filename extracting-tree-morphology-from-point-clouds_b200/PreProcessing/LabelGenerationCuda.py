"""Drop-in for the reference's ``PreProcessing/LabelGenerationCuda.py`` (variant A of the kernel:
perpendicularity tolerance 1e-6, no epsilon guards).  Same callables, same arguments, same outputs;
the work is done by the sm_100a library behind ``treemorph_b200.api``.

    closest_cylinder_cuda_batch          reference :20-111
    generate_offset_cloud_cuda_batched   reference :113-135
    label_clouds + CLI                   reference :137-234
"""
from __future__ import annotations

import argparse
import os
import re

import numpy as np
import pandas as pd

from .. import api, dropin
from ..Modules.Features import add_features
from ..Modules.Utils import get_device

VARIANT = api.VARIANT_A


def closest_cylinder_cuda_batch(points, start, radius, axis_length, axis_unit, IDs, device):
    return dropin.closest_cylinder(points, start, radius, axis_length, axis_unit, IDs, device, VARIANT)


def generate_offset_cloud_cuda_batched(cloud, cylinders, device, masterBar=None, batch_size=1024):
    return dropin.offset_cloud(cloud, cylinders, device, VARIANT, masterBar=masterBar, batch_size=batch_size)


def _numeric_prefix(path):
    head = os.path.basename(path).split(".")[0].split("_")
    return int(head[0]), int(head[1])


def _strip_to_digits(paths):
    """``--clean_data``: keep only digits and underscores of each base name (reference :145-167)."""
    for path in paths:
        folder, name = os.path.split(path)
        stem, ext = os.path.splitext(name)
        wanted = re.sub(r"[^\d_]", "", stem) + ext
        if wanted != name:
            target = os.path.join(folder, wanted)
            os.rename(path, target)
            print(f"Renamed: {path} -> {target}")


def _listing(folder, suffix):
    return [os.path.join(folder, f) for f in os.listdir(folder) if f.endswith(suffix)]


def label_clouds(cloudDir, cylinderDir, labelDir, batch_size=1024, clean_data=False, use_features=True):
    device = get_device()
    if clean_data:
        _strip_to_digits(_listing(cloudDir, ".npy"))
        _strip_to_digits(_listing(cylinderDir, ".csv"))
    clouds = sorted(_listing(cloudDir, ".npy"), key=_numeric_prefix)
    tables = sorted(_listing(cylinderDir, ".csv"), key=_numeric_prefix)

    print("\nLabeling clouds...")
    for cloud_path, table_path in zip(clouds, tables):
        cloud = np.load(cloud_path)
        cylinders = pd.read_csv(table_path, header=0)
        cylinders.columns = cylinders.columns.str.strip()
        stem = os.path.basename(cloud_path).split(".")[0]
        target = os.path.join(labelDir, stem + "_labeled.npy")
        if use_features:
            labelled = generate_offset_cloud_cuda_batched(cloud, cylinders, device, batch_size=batch_size)
            labelled = add_features(labelled, use_densities=False, use_curvatures=False, use_distances=False,
                                    use_verticalities=False)
            np.save(target, labelled)
        else:       # four dummy feature columns keep the (N,11) layout TreeSet expects: rows written once, into the file
            dropin.offset_cloud_to_npy(target, cloud, cylinders, device, VARIANT)
    print("Finished labeling and saving!")


def main(argv=None):
    parser = argparse.ArgumentParser(description="Label point clouds using cylinder models.")
    parser.add_argument("--cylinderDir", type=str, default=os.path.join("data", "raw", "QSM", "detailed"),
                        help="Directory containing the QSM cylinder CSV files.")
    parser.add_argument("--cloudDir", type=str, default=os.path.join("data", "raw", "cloud"),
                        help="Directory containing the raw point clouds.")
    parser.add_argument("--labelDir", type=str, default=os.path.join("data", "labeled", "offset", "cloud"),
                        help="Directory where labeled clouds will be saved.")
    parser.add_argument("--clean_data", action="store_true", help="Enable data cleaning before processing.")
    parser.add_argument("--use_features", action="store_true", help="Use additional features for labeling.")
    args = parser.parse_args(argv)
    here = os.getcwd()
    label_clouds(os.path.join(here, args.cloudDir), os.path.join(here, args.cylinderDir),
                 os.path.join(here, args.labelDir), clean_data=args.clean_data, use_features=args.use_features)


if __name__ == "__main__":
    main()
