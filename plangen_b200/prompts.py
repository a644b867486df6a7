"""Host prompt pipeline of PlanGen's inference modes, mirroring the reference's own
functions with a pluggable tokenizer (no vocabulary files are available offline; any object with
`encode(str) -> list[int]` and `decode(list[int]) -> str` works, e.g. the HF tokenizer of the Janus checkpoint):

  sft_prompt               VLChatProcessor.apply_sft_template_for_multi_turn_prompts (processing_vlm.py:137-177) with the
                           "deepseek" conversation template (janus/utils/conversation.py:76-91, :293-309)
  wrap_t2i_prompt          System.wrap_t2i_prompt     (plangen_base.py:210-229)
  wrap_uni_prompt          System.wrap_uni_prompt     (plangen_base.py:231-261)
  pad_input_ids            System.pad_input_ids       (plangen_base.py:699-725, test branch: LEFT padding)
  uni_batch / stage1_batch the `uni` and `uni_stage1` parts of System.mmu_collate (plangen_base.py:781-805)
  t2i_infer_collate_batch  System.t2i_infer_collate_batch (plangen_base.py:636-697), shared or per-sample negatives
  decode_plan_text_batch   System.decode_plan_text_batch  (plangen_base.py:296-306)
  plan_then_generate       the `uni_2stage` flow of System.test_step (plangen_base.py:369-400): stage-1 layout text
                           (x2t) -> re-wrapped prompt -> CFG image decode (t2i)
  mmu_process_one / mmu_batchify / mmu_infer_batch
                           the `mmu` / `mmu_infer` parts of System.mmu_collate (plangen_base.py:807-841) on top of
                           VLChatProcessor.process_one / add_image_token / batchify (processing_vlm.py:215-258, :260-324,
                           :361-423): `<image_placeholder>` expanded to boi + 576 image slots + eoi, LEFT padding,
                           images_seq_mask / images_emb_mask - the inputs of vl_gpt.prepare_inputs_embeds
  describe_then_ground     the `mmu` flow (plangen_base.py:851-881): prepare_inputs_embeds -> language_model.generate
  write_png / save_images  the PNG side of the result writers (plangen_base.py:444-453, :1162-1181): denorm_pt +
                           to_pil(...).save(...) without PIL (zlib + struct)

Pure host code (lists, strings, small int tensors): it defines the row order and mask contract the device path
consumes; no arithmetic lives here."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

import struct
import zlib

USER, ASSISTANT = "<|User|>", "<|Assistant|>"
IMAGE_PLACEHOLDER = "<image_placeholder>"       # processing_vlm.py:88 (image_tag)
# plangen_base.py:811, :828
MMU_QUESTION = "Please describe this image and then give the description and bounding box of each object in the image."
SEP, SEP2 = "\n\n", "<｜end▁of▁sentence｜>"
IMAGE_START_TAG = "<begin_of_image>"            # processing_vlm.py:89
# cfg/base.py:129
DEFAULT_NEG_PROMPT = ("low quality, jpeg artifacts, ugly, duplicate, morbid, mutilated, extra fingers, mutated hands, "
                      "poorly drawn hands, poorly drawn face, mutation, deformed, blurry, dehydrated, bad anatomy, bad "
                      "proportions, extra limbs, cloned face, disfigured, gross proportions, malformed limbs, missing arms, "
                      "missing legs, extra arms, extra legs, fused fingers, too many fingers.")


def sft_prompt(conversation: Sequence[Dict[str, str]], system_prompt: str = "") -> str:
    """The "deepseek" SFT template: `role: content` turns, user turns closed by a blank line and assistant turns by the
    end-of-sentence tag, an empty turn rendered as `role:`; contents and the result are stripped."""
    out = (system_prompt + SEP) if system_prompt else ""
    for i, msg in enumerate(conversation):
        content = msg["content"].strip()
        if content:
            out += msg["role"] + ": " + content + (SEP if i % 2 == 0 else SEP2)
        else:
            out += msg["role"] + ":"
    return out.strip()


class PromptPipeline:
    def __init__(self, tokenizer, pad_id: int, image_token_num: int = 576, neg_prompt: str = DEFAULT_NEG_PROMPT,
                 image_id: Optional[int] = None, image_start_id: Optional[int] = None, image_end_id: Optional[int] = None):
        """`image_id` / `image_start_id` / `image_end_id`: vocabulary ids of `<image_placeholder>`, `<begin_of_image>`,
        `<end_of_image>` (VLChatProcessor.image_id / image_start_id / image_end_id, processing_vlm.py:179-197); only the
        mmu functions need them."""
        self.tok, self.pad_id, self.n_img, self.neg_prompt = tokenizer, int(pad_id), int(image_token_num), neg_prompt
        self.image_id, self.image_start_id, self.image_end_id = image_id, image_start_id, image_end_id

    # ------------------------------------------------------------------ single prompts
    def wrap_t2i_prompt(self, caption: str) -> Tuple[str, torch.Tensor]:
        prompt = sft_prompt([{"role": USER, "content": caption}, {"role": ASSISTANT, "content": ""}]) + IMAGE_START_TAG
        return prompt, torch.LongTensor(self.tok.encode(prompt))

    def wrap_uni_prompt(self, caption: str, grounding: Optional[str] = None, in_stage1: bool = False) -> Tuple[str, torch.Tensor]:
        prompt = sft_prompt([{"role": USER, "content": caption}, {"role": ASSISTANT, "content": f"{grounding}"}])
        if not in_stage1:
            prompt = prompt + IMAGE_START_TAG
        ids = torch.LongTensor(self.tok.encode(prompt))
        if in_stage1:
            ids = ids[..., :-1]          # drop the closing tag: the model continues the assistant turn
        return prompt, ids

    # ------------------------------------------------------------------ batches
    def pad_input_ids(self, all_ids: Sequence[Sequence[int]], max_length: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        bs = len(all_ids)
        if max_length is None:
            max_length = max(len(t) for t in all_ids)
        ids = torch.full((bs, max_length), self.pad_id, dtype=torch.long)
        mask = torch.zeros((bs, max_length), dtype=torch.long)
        for i, t in enumerate(all_ids):
            n = len(t)
            if n > max_length:
                raise ValueError("prompt longer than max_length")
            if n:
                ids[i, max_length - n:] = torch.as_tensor(list(t), dtype=torch.long)
                mask[i, max_length - n:] = 1
        return ids, mask

    def uni_batch(self, base_captions: Sequence[str], groundings: Sequence[str]) -> Tuple[torch.Tensor, torch.Tensor]:
        ids, mask = self.pad_input_ids([self.wrap_uni_prompt(c, g)[1] for c, g in zip(base_captions, groundings)])
        return ids, torch.cat([mask, torch.ones((len(base_captions), self.n_img), dtype=torch.long)], dim=-1)

    def stage1_batch(self, base_captions: Sequence[str]) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.pad_input_ids([self.wrap_uni_prompt(c, "<grounding>", in_stage1=True)[1] for c in base_captions])

    def t2i_infer_collate_batch(self, uni_ids: torch.Tensor, uni_mask: torch.Tensor,
                                neg: Optional[Tuple[Sequence[str], Sequence[str]]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Interleaved cond / negative rows: ids (2B, P) int32, mask (2B, P + n_img) int32.  `neg` = per-sample
        (captions, groundings) (`use_neg_box`), else the shared negative prompt with an empty grounding."""
        bs, max_length = uni_ids.shape
        if neg is not None:
            neg_all = [self.wrap_uni_prompt(c, g)[1] for c, g in zip(*neg)]
        else:
            neg_all = [self.wrap_uni_prompt(self.neg_prompt, "")[1]] * bs
        max_neg = max(len(t) for t in neg_all)
        if max_neg > max_length:                      # negatives longer than every cond prompt: pad the cond side on the left
            need = max_neg - max_length
            uni_ids = torch.cat([torch.full((bs, need), self.pad_id, dtype=uni_ids.dtype), uni_ids], dim=1)
            uni_mask = torch.cat([torch.zeros((bs, need), dtype=uni_mask.dtype), uni_mask], dim=1)
            max_length = max_neg
        neg_ids, neg_mask = self.pad_input_ids(neg_all, max_length=max_length)
        neg_mask = torch.cat([neg_mask, torch.ones((bs, self.n_img), dtype=torch.long)], dim=-1)
        ids = torch.stack([uni_ids.long(), neg_ids], dim=1).view(bs * 2, -1)
        mask = torch.stack([uni_mask.long(), neg_mask], dim=1).view(bs * 2, -1)
        return ids.int(), mask.int()

    # ------------------------------------------------------------------ stage-1 output
    def decode_plan_text_batch(self, token_rows: Sequence[Sequence[int]]) -> List[str]:
        out = []
        for row in token_rows:
            text = "<grounding>" + self.tok.decode([int(t) for t in row])
            end = text.find("</grounding>")
            out.append(text[:end + len("</grounding>")] if end != -1 else "<grounding>" + "</grounding>")
        return out

    # ------------------------------------------------------------------ uni_2stage
    def plan_then_generate(self, engine, base_captions: Sequence[str], eos_token_id: int, bos_token_id: Optional[int] = None,
                           max_new_tokens: int = 512, cfg_weight: float = 5.0, temperature: float = 1.0):
        """Layout-image joint generation: the model first writes the layout (`<grounding>...</grounding>`), the prompt
        is re-wrapped with it, then the image tokens are decoded under CFG.  Returns (images, layout texts)."""
        s1_ids, s1_mask = self.stage1_batch(base_captions)
        emb = engine.language_model.get_input_embeddings()(s1_ids.to(engine.device))
        new = engine.language_model.generate(inputs_embeds=emb, attention_mask=s1_mask.to(engine.device), pad_token_id=eos_token_id,
                                             bos_token_id=bos_token_id, eos_token_id=eos_token_id, max_new_tokens=max_new_tokens,
                                             do_sample=False, use_cache=True)
        layouts = self.decode_plan_text_batch(new.cpu().tolist())
        uni_ids, uni_mask = self.uni_batch(base_captions, layouts)
        ids, mask = self.t2i_infer_collate_batch(uni_ids, uni_mask)
        dec, _ = engine.t2i(tokens=ids.to(engine.device), mask=mask.to(engine.device), cfg_weight=cfg_weight, temperature=temperature,
                            image_token_num_per_image=self.n_img)
        return dec, layouts


    # ------------------------------------------------------------------ mmu (image understanding)
    def mmu_process_one(self, images: torch.Tensor, answer: str = "", question: str = MMU_QUESTION) -> Dict[str, object]:
        """VLChatProcessor.process_one on the conversation System.mmu_collate builds (plangen_base.py:812-822, :829-839):
        user turn `<image_placeholder>\n{question}`, assistant turn `answer` (empty for inference).  `images`:
        (n, 3, H, W) pixel tensors already in the vision tower's input range.  Every placeholder token becomes
        [boi, image_id x n_img, eoi] (add_image_token, processing_vlm.py:215-258, add_special_token = False)."""
        if self.image_id is None or self.image_start_id is None or self.image_end_id is None:
            raise ValueError("mmu prompts need image_id / image_start_id / image_end_id")
        text = sft_prompt([{"role": USER, "content": f"{IMAGE_PLACEHOLDER}\n{question}"}, {"role": ASSISTANT, "content": f"{answer}"}])
        ids = torch.LongTensor(self.tok.encode(text))
        where = (ids == self.image_id).nonzero().flatten().tolist()
        pieces, start = [], 0
        for idx in where:
            pieces += [ids[start:idx], torch.tensor([self.image_start_id]), torch.full((self.n_img,), self.image_id, dtype=torch.long),
                       torch.tensor([self.image_end_id])]
            start = idx + 1
        pieces.append(ids[start:])
        return {"sft_format": text, "input_ids": torch.cat(pieces), "pixel_values": images, "num_image_tokens": [self.n_img] * len(where)}

    def mmu_batchify(self, prepares: Sequence[Dict[str, object]]) -> Dict[str, torch.Tensor]:
        """VLChatProcessor.batchify (processing_vlm.py:361-423): LEFT padding with pad_id, attention mask, pixel values
        stacked to (b, max_n_images, 3, H, W), images_seq_mask (b, T) = image slots, images_emb_mask (b, max_n, n_img)."""
        bs = len(prepares)
        T = max(len(p["input_ids"]) for p in prepares)
        max_n = max(1, max(len(p["num_image_tokens"]) for p in prepares))
        shape = next((tuple(p["pixel_values"].shape[1:]) for p in prepares if len(p["num_image_tokens"])), (3, 384, 384))
        out = {"input_ids": torch.full((bs, T), self.pad_id, dtype=torch.long), "attention_mask": torch.zeros((bs, T), dtype=torch.long),
               "pixel_values": torch.zeros((bs, max_n) + shape), "images_seq_mask": torch.zeros((bs, T), dtype=torch.bool),
               "images_emb_mask": torch.zeros((bs, max_n, self.n_img), dtype=torch.bool), "sft_format": [p["sft_format"] for p in prepares]}
        for i, p in enumerate(prepares):
            ids, n = p["input_ids"], len(p["input_ids"])
            out["attention_mask"][i, T - n:] = 1
            out["input_ids"][i, T - n:] = ids
            out["images_seq_mask"][i, T - n:] = ids == self.image_id
            k = len(p["num_image_tokens"])
            if k:
                out["pixel_values"][i, :k] = p["pixel_values"]
                for j, cnt in enumerate(p["num_image_tokens"]):
                    out["images_emb_mask"][i, j, :cnt] = True
        return out

    def mmu_infer_batch(self, images: torch.Tensor, answers: Optional[Sequence[str]] = None) -> Dict[str, torch.Tensor]:
        """`prepare_inputs_infer` (answers None: empty assistant turn) / `prepare_inputs` of System.mmu_collate for a batch of
        images (b, 3, H, W), one image per sample."""
        return self.mmu_batchify([self.mmu_process_one(images[i:i + 1], "" if answers is None else answers[i]) for i in range(len(images))])

    def describe_then_ground(self, engine, images: torch.Tensor, eos_token_id: int, bos_token_id: Optional[int] = None,
                             max_new_tokens: int = 512) -> List[str]:
        """Image layout understanding (`mmu`, plangen_base.py:851-881): SigLIP + aligner + scatter (prepare_inputs_embeds),
        then greedy decode of the description / boxes text."""
        b = self.mmu_infer_batch(images)
        x = engine.prepare_inputs_embeds(input_ids=b["input_ids"], pixel_values=b["pixel_values"], images_seq_mask=b["images_seq_mask"],
                                         images_emb_mask=b["images_emb_mask"])
        new = engine.language_model.generate(inputs_embeds=x, attention_mask=b["attention_mask"].to(engine.device), pad_token_id=eos_token_id,
                                             bos_token_id=bos_token_id, eos_token_id=eos_token_id, max_new_tokens=max_new_tokens,
                                             do_sample=False, use_cache=True)
        return [self.tok.decode([int(t) for t in row if int(t) != eos_token_id]) for row in new.cpu().tolist()]


# ---------------------------------------------------------------------- image writer (no PIL needed)
def write_png(path: str, img) -> None:
    """8-bit RGB (H, W, 3) or grey (H, W) array / tensor -> PNG file (zlib-compressed, filter 0 on every scanline)."""
    a = img.detach().cpu().numpy() if isinstance(img, torch.Tensor) else img
    if a.dtype.name != "uint8" or a.ndim not in (2, 3) or (a.ndim == 3 and a.shape[2] != 3):
        raise ValueError("write_png takes uint8 (H, W, 3) or (H, W)")
    h, w = a.shape[:2]
    raw = b"".join(b"\x00" + a[y].tobytes() for y in range(h))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2 if a.ndim == 3 else 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def save_images(engine, dec: torch.Tensor, path_format: str, start_index: int = 0) -> List[str]:
    """`to_pil(denorm_pt(pr_image[i])).save(f"{path}/pr_image/{idx}.png")` (plangen_base.py:1174-1177): (B, 3, H, W) decoder
    output -> one PNG per image; denorm + uint8 conversion on the device (engine.images_to_uint8), files named
    path_format.format(index)."""
    u8 = engine.images_to_uint8(dec).permute(0, 2, 3, 1).contiguous().cpu()
    paths = []
    for i in range(u8.shape[0]):
        paths.append(path_format.format(start_index + i))
        write_png(paths[-1], u8[i])
    return paths
