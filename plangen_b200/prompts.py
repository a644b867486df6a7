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

Pure host code (lists, strings, small int tensors): it defines the row order and mask contract the device path
consumes; no arithmetic lives here."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

USER, ASSISTANT = "<|User|>", "<|Assistant|>"
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
    def __init__(self, tokenizer, pad_id: int, image_token_num: int = 576, neg_prompt: str = DEFAULT_NEG_PROMPT):
        self.tok, self.pad_id, self.n_img, self.neg_prompt = tokenizer, int(pad_id), int(image_token_num), neg_prompt

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
