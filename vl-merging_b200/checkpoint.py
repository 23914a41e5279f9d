"""checkpoint.py — the step right before the merge: turning a loaded VLMo checkpoint
into the state_dict the merge methods consume.  Mirrors ViLTransformerSS.modify_checkpoint_vlmo
(src/vilt/modules/vilt_module.py:749-806) as a free function of (ckpt, config):

  * accepts a PL checkpoint ({"state_dict": ...}) or a bare state_dict (:751-755);
  * if the text position table is longer than max_text_len, truncates it (and position_ids) (:757-763);
  * drops the index buffers that are rebuilt from the config (:778-784);
  * if the checkpoint was trained at another image size, bicubically resizes the image part of
    relative_position_bias_table ((2W-1)^2 rows x heads*layers) and keeps the text / extra rows (:786-804).

Host-side plumbing in stock torch (the table is a few hundred KB); like the reference it edits the dict
in place and returns it.
"""
import torch

POP_KEYS = ("relative_position_index", "text_relative_position_index", "text_imag_relative_position_index",
            "video_relative_position_index", "text_video_relative_position_index", "temporal_relative_position_index",
            "mask_for_combining_temporal")


def load_checkpoint(path):
    """torch.load of a PL .ckpt / bare state_dict saved by the reference (pickled, map_location cpu)."""
    return torch.load(path, map_location="cpu", weights_only=False)


def save_checkpoint(state_dict, path):
    """Writes {"state_dict": ...}, the layout the reference's loaders accept (vilt_module.py:751-755, :354)."""
    torch.save({"state_dict": dict(state_dict)}, path)


def modify_checkpoint_vlmo(ckpt, config):
    state_dict = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    max_text_len = config["max_text_len"]
    pos = "text_embeddings.position_embeddings.weight"
    if state_dict[pos].size(0) != max_text_len:
        state_dict[pos] = state_dict[pos][:max_text_len, :]
        if "text_embeddings.position_ids" in state_dict:  # not persistent in recent transformers versions
            state_dict["text_embeddings.position_ids"] = state_dict["text_embeddings.position_ids"][:, :max_text_len]
        for k in POP_KEYS[:3]:
            state_dict.pop(k, None)

    table = state_dict["relative_position_bias_table"]
    src_num_pos = table.size(0)
    w = config["image_size"] // config["patch_size"]
    text_rows = 2 * config["max_text_len_of_initckpt"]
    dst_num_pos = (2 * w - 1) * (2 * w - 1) + 3 + text_rows + 2
    non_image = text_rows + 2 + 3  # text distances + text extras + image extras (:773)
    src_size = int((src_num_pos - non_image) ** 0.5)
    dst_size = int((dst_num_pos - non_image) ** 0.5)
    for k in POP_KEYS:
        state_dict.pop(k, None)
    if src_size != dst_size:
        extra = table[-non_image:, :]
        image = table[:-non_image, :]
        embed = image.transpose(0, 1).reshape(-1, src_size, src_size)
        embed = torch.nn.functional.interpolate(embed.unsqueeze(0), size=(dst_size, dst_size), mode="bicubic")
        embed = embed.squeeze(0).permute(1, 2, 0).contiguous().view(-1, embed.size(1))
        state_dict["relative_position_bias_table"] = torch.cat((embed, extra), dim=0)
    return state_dict
