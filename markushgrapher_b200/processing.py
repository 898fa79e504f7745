"""Host-side mirrors of the fork's image processor / tokenizer / processor classes (reference begin.py:105-111,
utils/common.py:34-42) and of the OCR-cell packing in front of them (core/common/data_preprocessing.py:24-104,
core/datasets/task_collator.py:28-107). The tensors they emit are the hot path's input contract:
input_ids (1,Lt) i64, bbox (1,Lt,4) f32 in [0,1], attention_mask (1,Lt), pixel_values (1,3,512,512) f32.
String / sub-word work stays on the host (it is sentencepiece-bound); the pixel side has a device implementation in
packing.py."""
from __future__ import annotations

import os
import re
from typing import List, Optional

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------ OCR cells
def split_bounding_box_for_words(sentence: str, bounding_box, tokenizer):
    """One OCR cell -> its sentencepiece pieces and one box per piece: the cell's width is shared out in proportion
    to a 12-px-per-character estimate, left to right (reference data_preprocessing.py:16-48).  Arithmetic is in
    Python floats (doubles) in the reference's order -- fraction first, then the running left edge -- so the boxes
    are bit-identical (tests/test_cells_cpu.py replays golden vectors from the reference's own function)."""
    pieces = tokenizer.tokenize(sentence)
    widths = np.array([12 * (1 if p == "▁" else sum(ch != "▁" for ch in p)) for p in pieces], dtype=np.float64)
    x_min, y_min, x_max, y_max = bounding_box
    if len(pieces) == 0:
        return pieces, []
    adjusted = (x_max - x_min) * (widths / float(widths.sum()))
    edges = np.cumsum(np.concatenate(([x_min], adjusted)))          # sequential: left_{i+1} = left_i + adjusted_i
    return pieces, [(float(edges[i]), y_min, float(edges[i] + adjusted[i]), y_max) for i in range(len(pieces))]


def prepare_cells_to_text(cells, tokenizer, w, h, normalize_bbox: bool, max_sequence_length: int = 512):
    """OCR cells [{"text", "bbox" in [0,1]}] -> (words, boxes, token count): the in-memory hand-off from ChemicalOCR
    (reference ocr/chemical_ocr.py:396-478 writes exactly these cells to disk) to the processor, reference
    data_preprocessing.py:59-104.  normalize_bbox=False is the pre-2025 0..500 integer box format."""
    words, boxes, n_tok = [], [], 0
    for cell in cells:
        text = cell["text"]
        if text.isspace():
            continue
        bx = cell["bbox"]
        pieces, piece_boxes = split_bounding_box_for_words(text, [bx[0] * w, bx[1] * h, bx[2] * w, bx[3] * h], tokenizer)
        for piece, box in zip(pieces, piece_boxes):
            if piece.isspace():
                continue
            if not normalize_bbox:
                box = (int((box[0] / w) * 500), int((box[1] / h) * 500), int((box[2] / w) * 500), int((box[3] / h) * 500))
            if max(box) > 500:
                continue
            word = str(piece).strip()
            words.append(word)
            boxes.append(box)
            n_tok += len(tokenizer.tokenize(word))
            if n_tok >= max_sequence_length - 15:
                break
        if n_tok >= max_sequence_length:
            break
    return words, boxes, n_tok


def collate_cells(image, cells, tokenizer, question: str, normalize_bbox: bool = True):
    """(page image, OCR cells, question) -> (image, instruction, words, boxes) as TaskCollator.collate returns them
    (reference core/datasets/task_collator.py:28-107; the label half is training-only)."""
    w, h = image.size
    words, boxes, _ = prepare_cells_to_text(cells, tokenizer, w, h, normalize_bbox)
    if normalize_bbox:
        boxes = [[b[0] / w, b[1] / h, b[2] / w, b[3] / h] for b in boxes]
    return image, f"Question Answering. {question}", words, boxes


# ------------------------------------------------------------------------------------------------ ChemicalOCR hand-off
_LOC = re.compile(r"<loc_(\d+)>")
_LOC4 = re.compile(r"(?:<loc_\d+>){4}")
_NUM_LINE = re.compile(r"^(?:\d+>)*(\d+)>(\d+)>(\d+)>(\d+)>(.+)$")


def clean_ocr_text(text: str, start_tag: str = "<ocr>", end_tag: Optional[str] = "</ocr>") -> str:
    """keep what lies between the first `start_tag` and the first `end_tag`, tags included (the VLM may chat before and
    after its answer) -- reference ocr/chemical_ocr.py:202-223"""
    i = text.find(start_tag)
    if i >= 0:
        text = text[i:]
    if end_tag:
        j = text.find(end_tag)
        if j >= 0:  # (the reference's `.*?$` stops in front of a final newline, which therefore survives)
            text = text[: j + len(end_tag)] + ("\n" if text.endswith("\n") else "")
    return text


def parse_ocr_string(ocr_string: str):
    """ChemicalOCR's generated string -> (words, boxes normalised by the 500-unit page), reference
    ocr/chemical_ocr.py:165-199.  Two line formats: legacy `<loc_x1><loc_y1><loc_x2><loc_y2>text` (after an optional
    leading page box `<loc_0><loc_0><loc_500><loc_500>`), and `x1>y1>x2>y2>text` with any number of leading `N>`
    fields skipped (the page box on the first line).  Lines without text or with fewer than four numbers are dropped;
    the LAST four numbers of a legacy line are its box."""
    body = ocr_string.replace("<ocr>", "").replace("</ocr>", "").strip()
    words, boxes = [], []
    if "<loc_" in body:
        page = "<loc_0><loc_0><loc_500><loc_500>"
        if body.startswith(page):
            body = body[len(page):].strip()
        for line in body.splitlines():
            nums = [int(n) for n in _LOC.findall(line)]
            text = _LOC4.sub("", line).strip()
            if len(nums) >= 4 and text:
                words.append(text)
                boxes.append([v / 500 for v in nums[-4:]])
    else:
        for line in body.splitlines():
            m = _NUM_LINE.match(line.strip())
            if not m:
                continue
            text = m.group(5).strip()
            if text:
                words.append(text)
                boxes.append([int(m.group(k)) / 500 for k in (1, 2, 3, 4)])
    return words, boxes


def cells_from_ocr_string(output_text: str):
    """one generated string -> the OCR cells [{"bbox", "text"}] the reference stores per image
    (ocr/chemical_ocr.py:447-455); feed them to MarkushgrapherProcessor.from_cells"""
    words, boxes = parse_ocr_string(clean_ocr_text(output_text))
    return [{"bbox": b, "text": w} for w, b in zip(words, boxes)]


class MarkushgrapherImageProcessor:
    def __init__(self, apply_ocr: bool = False, size=None, image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5), **kw):
        if apply_ocr:
            raise ValueError("apply_ocr=True (pytesseract) is not part of the hot path; OCR comes from ChemicalOCR")
        size = size or {"height": 512, "width": 512}
        self.size = (int(size["height"]), int(size["width"]))
        self.mean = torch.tensor(image_mean).view(3, 1, 1)
        self.std = torch.tensor(image_std).view(3, 1, 1)

    def __call__(self, images, return_tensors="pt", **kw):
        from PIL import Image

        if not isinstance(images, (list, tuple)):
            images = [images]
        out = []
        for im in images:
            im = im.convert("RGB").resize((self.size[1], self.size[0]), resample=Image.BILINEAR)
            t = torch.from_numpy(np.asarray(im, dtype=np.uint8).copy()).permute(2, 0, 1).float() * (1.0 / 255.0)
            out.append((t - self.mean) / self.std)
        return {"pixel_values": torch.stack(out)}


class MarkushgrapherTokenizer:
    """SentencePiece (T5/UDOP vocabulary) wrapper with the handful of methods markushgrapher.core calls
    (markush_tokenizer.py:309-618, utils/common.py:47-64). EOS=1, PAD=0, vocab 33201."""

    def __init__(self, sp_model_path: str, vocab_size: int = 33201, extra_ids: int = 100, loc_extra_ids: int = 501,
                 other_extra_ids: int = 200):
        import sentencepiece as spm

        self.sp = spm.SentencePieceProcessor()
        self.sp.Load(sp_model_path)
        self.vocab_size = vocab_size
        self.eos_token_id, self.pad_token_id, self.unk_token_id = 1, 0, 2
        self.eos_token, self.pad_token = "</s>", "<pad>"
        n = self.sp.GetPieceSize()
        self.special = {}
        # UDOP appends <extra_id_*>, <extra_l_id_*>, </extra_l_id_*>, <extra_t_id_*>, </extra_t_id_*>, <loc_*>, <other_*>
        # after the sentencepiece vocabulary; ids are assigned in reverse order like the stock tokenizer
        names = ([f"<extra_id_{i}>" for i in range(extra_ids - 1, -1, -1)] +
                 [f"<extra_l_id_{i}>" for i in range(extra_ids - 1, -1, -1)] +
                 [f"</extra_l_id_{i}>" for i in range(extra_ids - 1, -1, -1)] +
                 [f"<extra_t_id_{i}>" for i in range(extra_ids - 1, -1, -1)] +
                 [f"</extra_t_id_{i}>" for i in range(extra_ids - 1, -1, -1)] +
                 [f"<loc_{i}>" for i in range(loc_extra_ids - 1, -1, -1)] +
                 [f"<other_{i}>" for i in range(other_extra_ids - 1, -1, -1)])
        for i, s in enumerate(names):
            self.special[s] = n + i
        self.special_inv = {v: k for k, v in self.special.items()}

    @classmethod
    def from_pretrained(cls, path: str, **kw):
        fn = os.path.join(path, "spiece.model") if os.path.isdir(path) else path
        if not os.path.exists(fn):
            raise FileNotFoundError(f"{fn} not found: MarkushgrapherTokenizer needs the UDOP sentencepiece model")
        return cls(fn, **kw)

    def __len__(self):
        return self.vocab_size

    def tokenize(self, text: str) -> List[str]:
        return self.sp.EncodeAsPieces(text)

    def _convert_token_to_id(self, token: str) -> int:
        return self.special.get(token, self.sp.PieceToId(token))

    def convert_tokens_to_ids(self, tokens):
        if isinstance(tokens, str):
            return self._convert_token_to_id(tokens)
        return [self._convert_token_to_id(t) for t in tokens]

    def convert_ids_to_tokens(self, ids):
        if isinstance(ids, int):
            return self.special_inv.get(ids) or self.sp.IdToPiece(ids)
        return [self.convert_ids_to_tokens(int(i)) for i in ids]

    def encode(self, text: str, add_special_tokens: bool = True, **kw) -> List[int]:
        ids = self.convert_tokens_to_ids(self.tokenize(text))
        return ids + [self.eos_token_id] if add_special_tokens else ids

    def decode(self, ids, skip_special_tokens: bool = False, **kw) -> str:
        ids = [int(i) for i in (ids.tolist() if hasattr(ids, "tolist") else ids)]
        out, cur = [], []
        for i in ids:
            if i in self.special_inv or i in (0, 1):
                if cur:
                    out.append(self.sp.DecodeIds(cur))
                    cur = []
                if not skip_special_tokens:
                    out.append(self.special_inv.get(i, "</s>" if i == 1 else "<pad>"))
            else:
                cur.append(i)
        if cur:
            out.append(self.sp.DecodeIds(cur))
        return "".join(out)


class MarkushgrapherProcessor:
    """processor(images=PIL, text=[prompt], text_pair=[[word,...]], boxes=[[[x0,y0,x1,y1],...]], return_tensors="pt")
    -> input_ids, bbox, attention_mask, pixel_values with leading batch dim 1 (reference utils/common.py:34-42,68-71).
    Prompt tokens and separators get box `sep_box`/zeros like the UDOP tokenizer (boxes are already normalised to
    [0,1] by the reference's data pipeline, `normalize_bbox: True` in predict.yaml)."""

    def __init__(self, image_processor: MarkushgrapherImageProcessor, tokenizer: MarkushgrapherTokenizer,
                 sep_box=(1.0, 1.0, 1.0, 1.0)):
        self.image_processor = image_processor
        self.tokenizer = tokenizer
        self.sep_box = list(sep_box)

    def from_cells(self, image, cells, question: str = "What markush structure is in the image?",
                   normalize_bbox: bool = True):
        """ChemicalOCR cells straight to model inputs (no dataset round trip through the disk): collate + __call__,
        the two steps of reference utils/common.py:14-42"""
        image, instruction, words, boxes = collate_cells(image, cells, self.tokenizer, question, normalize_bbox)
        return self(images=image.convert("RGB"), text=[instruction], text_pair=[words], boxes=[boxes])

    def __call__(self, images=None, text=None, text_pair=None, boxes=None, return_tensors="pt", padding=False,
                 truncation=False, max_length: Optional[int] = None, **kw):
        tk = self.tokenizer
        ids, bb = [], []
        prompt = text[0] if isinstance(text, (list, tuple)) else text
        for t in tk.encode(prompt, add_special_tokens=False):
            ids.append(t)
            bb.append([0.0, 0.0, 0.0, 0.0])
        ids.append(tk.eos_token_id)
        bb.append(self.sep_box)
        words = text_pair[0] if text_pair and isinstance(text_pair[0], (list, tuple)) else (text_pair or [])
        wboxes = boxes[0] if boxes and boxes[0] and isinstance(boxes[0][0], (list, tuple)) else (boxes or [])
        for w, b in zip(words, wboxes):
            for t in tk.encode(w, add_special_tokens=False):
                ids.append(t)
                bb.append([float(v) for v in b])
        ids.append(tk.eos_token_id)
        bb.append(self.sep_box)
        if truncation and max_length:
            ids, bb = ids[:max_length], bb[:max_length]
        out = {"input_ids": torch.tensor([ids], dtype=torch.long), "bbox": torch.tensor([bb], dtype=torch.float32),
               "attention_mask": torch.ones(1, len(ids), dtype=torch.long)}
        if images is not None:
            out.update(self.image_processor(images))
        return out
