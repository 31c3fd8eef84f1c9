"""CPU restatement of the stand-alone NavDP policy's inference path -- TEST INFRASTRUCTURE.

Follows internnav/model/basemodel/navdp/navdp_policy.py (`NavDPNet`: predict_noise L161-170, predict_critic L172-187,
predict_pointgoal_batch_action_vel L302-322, predict_nogoal_batch_action_vel L324-339) and
internnav/model/encoder/navdp_backbone.py (`RGBDBackbone` L205-283, `LearnablePositionalEncoding` L40-57), built from the
primitives of oracle/navdp_oracle.py (DINOv2 ViT-S, nn.MultiheadAttention, post-norm decoder layer).  Plain fp32 PyTorch;
`sd` holds the reference's state_dict names.  Pinned by tests/golden/navdp_policy_reference.npz = outputs of the
reference's own NavDPNet run in this container (oracle/gen_golden_navdp_policy.py); per-environment semantics for B > 1
(sample i is conditioned on environment i // Ns; for B = 1 exactly the reference's `repeat`).
"""
import torch
import torch.nn.functional as F

from . import ddpm
from . import navdp_oracle as O


def rgbd_backbone(sd, images, depths, memory_size=8, p="rgbd_encoder."):
    """RGBDBackbone.forward, 5-D inputs: images [B, m, 224, 224, 3] in [0, 1], depths [B, 1, 224, 224, 1] -> [B, 16 m, 384]."""
    dt = images.dtype
    B, T = images.shape[:2]
    mean = torch.tensor([0.485, 0.456, 0.406], dtype=torch.float32, device=images.device).reshape(1, 3, 1, 1).to(dt)
    std = torch.tensor([0.229, 0.224, 0.225], dtype=torch.float32, device=images.device).reshape(1, 3, 1, 1).to(dt)
    ti = images.permute(0, 1, 4, 2, 3).reshape(-1, 3, 224, 224)
    image_token = O.dinov2_vits(sd, p + "rgb_model.", (ti - mean) / std).reshape(B, T * 256, -1)
    td = depths.permute(0, 1, 4, 2, 3).reshape(-1, 1, 224, 224)
    Td = depths.shape[1]
    depth_token = O.dinov2_vits(sd, p + "depth_model.", torch.cat([td, td, td], dim=1)).reshape(B, Td * 256, -1)
    tok = torch.cat((image_token, depth_token), dim=1)
    tok = tok + sd[p + "former_pe.position_embedding.weight"][: tok.shape[1]].to(dt)
    x = sd[p + "former_query.position_embedding.weight"][: memory_size * 16].to(dt).unsqueeze(0).expand(B, -1, -1)
    for i in range(2):
        x = O._decoder_layer_post(sd, "%sformer_net.layers.%d." % (p, i), x, tok, 8)
    return O._lin(sd, p + "project_layer", x)


def _cond(sd, slots, rgbd_embed, Ns):
    """[slot_0 .. slot_3, memory tokens] + cond_pos_embed, repeated for the Ns samples of every environment."""
    c = torch.cat(list(slots) + [rgbd_embed], dim=1)
    c = c + sd["cond_pos_embed.position_embedding.weight"][: c.shape[1]].to(c.dtype)
    return c.repeat_interleave(Ns, dim=0)


def _decoder(sd, x, cond, causal, mem_first=0, layers=16, heads=8):
    for i in range(layers):
        p = "decoder.layers.%d." % i
        h = O._ln(sd, p + "norm1", x, 1e-5)
        x = x + O._mha(sd, p + "self_attn", h, h, h, heads, causal=causal)
        h = O._ln(sd, p + "norm2", x, 1e-5)
        mem = cond[:, mem_first:]          # memory_mask = -inf on the first `mem_first` keys == those keys removed
        x = x + O._mha(sd, p + "multihead_attn", h, mem, mem, heads)
        h = O._ln(sd, p + "norm3", x, 1e-5)
        x = x + O._lin(sd, p + "linear2", F.gelu(O._lin(sd, p + "linear1", h)))
    return O._ln(sd, "layernorm", x, 1e-5)


def predict_noise(sd, last_actions, timestep, goal_embed, rgbd_embed):
    """NavDPNet.predict_noise: last_actions [B*Ns, T, 3]; timestep int64 [1]; goal [B, 1, 384]; rgbd [B, 16 m, 384]."""
    dt = last_actions.dtype
    B = goal_embed.shape[0]
    Ns = last_actions.shape[0] // B
    T = last_actions.shape[1]
    x = O._lin(sd, "input_embed", last_actions)
    te = O.sinusoidal_pos_emb(timestep.reshape(1).expand(B).to(last_actions.device)).unsqueeze(1).to(dt)
    cond = _cond(sd, [te, goal_embed, goal_embed, goal_embed], rgbd_embed, Ns)
    x = x + sd["out_pos_embed.position_embedding.weight"][:T].to(dt)
    return O._lin(sd, "action_head", _decoder(sd, x, cond, causal=True))


def predict_critic(sd, predict_trajectory, rgbd_embed):
    """NavDPNet.predict_critic: trajectories [B*Ns, T, 3] -> [B*Ns]."""
    dt = predict_trajectory.dtype
    B = rgbd_embed.shape[0]
    Ns = predict_trajectory.shape[0] // B
    T = predict_trajectory.shape[1]
    zero = torch.zeros_like(rgbd_embed[:, 0:1])
    x = O._lin(sd, "input_embed", predict_trajectory) + sd["out_pos_embed.position_embedding.weight"][:T].to(dt)
    cond = _cond(sd, [zero, zero, zero, zero], rgbd_embed, Ns)
    out = _decoder(sd, x, cond, causal=False, mem_first=4)     # no tgt_mask in predict_critic; cond_critic_mask on 4 keys
    return O._lin(sd, "critic_head", out.mean(dim=1))[:, 0]


def sample_and_rank(sd, goal_embed, rgbd_embed, x_init, step_noise, K=10):
    sch = ddpm.DDPMScheduler(num_train_timesteps=K, beta_schedule="squaredcos_cap_v2", clip_sample=True,
                             prediction_type="epsilon")
    sch.set_timesteps(K)
    sch.noise_queue = [step_noise[i] for i in range(step_noise.shape[0])]
    x = x_init
    for k in sch.timesteps:
        eps = predict_noise(sd, x, k.unsqueeze(0), goal_embed, rgbd_embed)
        x = sch.step(model_output=eps, timestep=k, sample=x).prev_sample
    critic = predict_critic(sd, x, rgbd_embed)
    traj = torch.cumsum(x / 4.0, dim=1)
    return traj[critic.argsort()[0:8]], traj[(-critic).argsort()[0:8]], x, critic


def predict_pointgoal_batch_action_vel(sd, goal_point, images, depths, x_init, step_noise, K=10, memory_size=8):
    rgbd = rgbd_backbone(sd, images, depths, memory_size)
    goal = O._lin(sd, "point_encoder", goal_point).unsqueeze(1)
    return sample_and_rank(sd, goal, rgbd, x_init, step_noise, K)


def predict_nogoal_batch_action_vel(sd, images, depths, x_init, step_noise, K=10, memory_size=8):
    rgbd = rgbd_backbone(sd, images, depths, memory_size)
    return sample_and_rank(sd, torch.zeros_like(rgbd[:, 0:1]), rgbd, x_init, step_noise, K)
