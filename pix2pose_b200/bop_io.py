"""The two dataset-side helpers the recognition path needs from ``tools/bop_io.py`` of the reference, plus the minimal
``bop_toolkit_lib.inout.load_json`` stand-in (the bop_toolkit submodule of the reference is empty):

* ``get_model_params`` (bop_io.py:33-42): ``norm_factor.json`` entry -> the ``obj_param`` 6-vector ``pix2pose.__init__``
  takes (recognition.py:17-18; written by tools/2_1_ply_file_to_3d_coord_model.py:99);
* ``get_target_list`` (bop_io.py:9-31): BOP ``test_targets_*.json`` -> per image ``[scene_id, im_id, obj_ids, inst_counts]``,
  the ``inst_counts`` / ``obj_id_targets`` the evaluation loop consumes (tools/5_evaluation_bop_basic.py:245-262).
Host-side file handling only; nothing here touches the device."""
import json

import numpy as np


def load_json(path, keys_to_int=False):
    """bop_toolkit_lib.inout.load_json: JSON file -> dict, optionally with integer keys (norm_factor.json is keyed by obj id)."""
    def convert(x):
        return {int(k) if k.lstrip("-").isdigit() else k: v for k, v in x.items()}

    with open(path, "r") as f:
        return json.load(f, object_hook=convert) if keys_to_int else json.load(f)


def get_model_params(model_param):
    """bop_io.py:33-42."""
    obj_param = np.zeros((6))
    for i, k in enumerate(("x_scale", "y_scale", "z_scale", "x_ct", "y_ct", "z_ct")):
        obj_param[i] = model_param[k]
    return obj_param


def get_target_list(target_path_or_list):
    """bop_io.py:9-31 (accepts the path of a BOP targets file or the already loaded list)."""
    targets = load_json(target_path_or_list) if isinstance(target_path_or_list, str) else target_path_or_list
    prev_imid, prev_sid, target_list = -1, -1, []
    obj_ids, inst_counts = [], []
    for tgt in targets:
        im_id, inst_count, obj_id, scene_id = tgt["im_id"], tgt["inst_count"], tgt["obj_id"], tgt["scene_id"]
        if prev_imid != im_id or prev_sid != scene_id:
            if prev_imid != -1:
                target_list.append([prev_sid, prev_imid, obj_ids, inst_counts])
            obj_ids, inst_counts = [obj_id], [inst_count]
        else:
            obj_ids.append(obj_id)
            inst_counts.append(inst_count)
        prev_imid, prev_sid = im_id, scene_id
    target_list.append([prev_sid, prev_imid, obj_ids, inst_counts])
    return target_list
