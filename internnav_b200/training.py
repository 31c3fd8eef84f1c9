"""Host side of the training branch: the batch layout `InternVLAN1ForCausalLM.forward` consumes.

`collate_traj_batch` restates what the reference collator adds for dual-system samples
(internnav/dataset/internvla_n1_lerobot_dataset.py L1155-1183 `process_input_with_traj_tokens`, L1185-1277 `__call__`):
n_query TRAJ tokens appended to every sample (`t_s_pos` = their start), right padding with the pad id, attention_mask =
input_ids != pad, per-sample frame stacks padded to the longest by repeating the last frame, `video_frame_num` = the
true frame counts.  Everything is integer / copy work; nothing here touches the GPU.
"""
import torch

TRAJ_TOKEN_INDEX = 151667
IGNORE_INDEX = -100
PAD_TOKEN_ID = 151643


def process_input_with_traj_tokens(input_ids, labels, max_seq_len, traj_token_length=4):
    """L1155-1183: truncate to max_seq_len, remember the length as t_s_pos, append the TRAJ tokens to ids and labels."""
    traj = torch.full((traj_token_length,), TRAJ_TOKEN_INDEX, dtype=input_ids[0].dtype)
    out_ids, out_labels, t_s_pos = [], [], []
    for ids, lab in zip(input_ids, labels):
        ids, lab = ids[:max_seq_len], lab[:max_seq_len]
        t_s_pos.append(len(ids))
        out_ids.append(torch.cat([ids, traj]))
        out_labels.append(torch.cat([lab, traj.clone()]))
    return out_ids, out_labels, t_s_pos


def _pad_frames(x, n):
    if x.shape[0] == n:
        return x
    return torch.cat([x, x[-1:].expand(n - x.shape[0], *x.shape[1:])], dim=0)


def collate_traj_batch(instances, model_max_length=8192, pad_token_id=PAD_TOKEN_ID, traj_token_length=4):
    """instances: dicts with input_ids [1,S] / labels [1,S] / pixel_values [n,1176] / image_grid_thw [k,3] /
    traj_images [f,224,224,3] / traj_depths [f,224,224] / traj_poses [f,T,3]  ->  the batch dict of L1209-1277."""
    ids = [i["input_ids"].squeeze(0) for i in instances]
    labels = [i["labels"].squeeze(0) for i in instances]
    ids, labels, t_s_pos = process_input_with_traj_tokens(ids, labels, model_max_length - traj_token_length,
                                                         traj_token_length)
    input_ids = torch.nn.utils.rnn.pad_sequence(ids, batch_first=True, padding_value=pad_token_id)[:, :model_max_length]
    labels = torch.nn.utils.rnn.pad_sequence(labels, batch_first=True, padding_value=IGNORE_INDEX)[:, :model_max_length]
    batch = dict(input_ids=input_ids, labels=labels, attention_mask=input_ids.ne(pad_token_id), position_ids=None,
                 t_s_pos=t_s_pos)
    with_px = [i for i in instances if "pixel_values" in i]
    batch["pixel_values"] = torch.cat([i["pixel_values"] for i in with_px], dim=0) if with_px else None
    batch["image_grid_thw"] = torch.cat([i["image_grid_thw"] for i in with_px], dim=0) if with_px else None
    n = max(i["traj_images"].shape[0] for i in instances)
    batch["video_frame_num"] = torch.tensor([i["traj_images"].shape[0] for i in instances])
    for key in ("traj_images", "traj_depths", "traj_poses"):
        batch[key] = torch.stack([_pad_frames(i[key], n) for i in instances])
    return batch
