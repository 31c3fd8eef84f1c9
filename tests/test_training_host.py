"""Collator mirror of the training branch (row a13, host side) -- CPU only."""
import torch

from internnav_b200.training import IGNORE_INDEX, PAD_TOKEN_ID, TRAJ_TOKEN_INDEX, collate_traj_batch


def _inst(S, f, k):
    return dict(input_ids=torch.arange(10, 10 + S).view(1, S), labels=torch.full((1, S), IGNORE_INDEX),
                pixel_values=torch.full((4 * k, 1176), float(k)), image_grid_thw=torch.tensor([[1, 2, 2 * k]]),
                traj_images=torch.arange(f).view(f, 1, 1, 1).expand(f, 224, 224, 3).float(),
                traj_depths=torch.arange(f).view(f, 1, 1).expand(f, 224, 224).float(),
                traj_poses=torch.arange(f).view(f, 1, 1).expand(f, 32, 3).float())


def test_collate_layout():
    b = collate_traj_batch([_inst(7, 3, 1), _inst(12, 1, 2), _inst(9, 2, 3)])
    assert b["t_s_pos"] == [7, 12, 9]
    assert b["input_ids"].shape == (3, 16)
    for i, t in enumerate(b["t_s_pos"]):
        assert b["input_ids"][i, t:t + 4].tolist() == [TRAJ_TOKEN_INDEX] * 4
        assert b["labels"][i, t:t + 4].tolist() == [TRAJ_TOKEN_INDEX] * 4
        assert (b["input_ids"][i, t + 4:] == PAD_TOKEN_ID).all() and (b["labels"][i, t + 4:] == IGNORE_INDEX).all()
        assert b["attention_mask"][i].sum() == t + 4
    assert b["pixel_values"].shape == (4 + 8 + 12, 1176) and b["image_grid_thw"].tolist() == [[1, 2, 2], [1, 2, 4], [1, 2, 6]]
    assert b["video_frame_num"].tolist() == [3, 1, 2]
    assert b["traj_images"].shape == (3, 3, 224, 224, 3)
    assert b["traj_poses"][1, :, 0, 0].tolist() == [0.0, 0.0, 0.0]        # single frame repeated
    assert b["traj_depths"][2, :, 0, 0].tolist() == [0.0, 1.0, 1.0]       # last frame repeated
    assert b["position_ids"] is None


def test_collate_truncates_before_appending():
    b = collate_traj_batch([_inst(30, 1, 1)], model_max_length=20)
    assert b["t_s_pos"] == [16] and b["input_ids"].shape == (1, 20)
    assert b["input_ids"][0, 16:].tolist() == [TRAJ_TOKEN_INDEX] * 4
