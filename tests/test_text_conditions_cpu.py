"""The text-condition adapter end to end from a checkpoint directory: tiny seeded CLIP-L / CLIP-G
/ T5 models and tokenizers are SAVED in the diffusers layout (tokenizer*/, text_encoder*/),
loaded back by `load_text_encoders` (what the pipeline constructor does when
`pretrained_model_name_or_path` holds them) and must reproduce the in-memory encoders."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_loader_roundtrip_and_prompt_batches(tmp_path):
    import sentencepiece as spm
    import transformers
    from common import tiny_text_stack
    from dwm.pipelines.text_conditions import (flatten_prompts, load_text_encoders,
                                               text_conditions)
    tok, encs, clip21 = tiny_text_stack()
    root = str(tmp_path)
    assert load_text_encoders(True, root, torch.device("cpu"), {}) is None      # nothing there
    tok.save_pretrained(os.path.join(root, "tokenizer"))
    tok.save_pretrained(os.path.join(root, "tokenizer_2"))
    encs[0].save_pretrained(os.path.join(root, "text_encoder"))
    encs[1].save_pretrained(os.path.join(root, "text_encoder_2"))
    corpus = os.path.join(root, "corpus.txt")
    with open(corpus, "w") as f:
        f.write("\n".join(["the car drives on the road at night", "a red car and a blue truck",
                           "rainy day in the city"] * 30))
    spm.SentencePieceTrainer.train(
        input=corpus, model_prefix=os.path.join(root, "t5tiny"), vocab_size=36,
        model_type="unigram", pad_id=0, eos_id=1, unk_id=2, bos_id=-1, hard_vocab_limit=False)
    t5tok = transformers.T5TokenizerFast(vocab_file=os.path.join(root, "t5tiny.model"), extra_ids=0)
    t5tok.save_pretrained(os.path.join(root, "tokenizer_3"))
    torch.manual_seed(3)
    t5 = transformers.T5EncoderModel(transformers.T5Config(
        vocab_size=64, d_model=96, d_kv=8, d_ff=64, num_layers=2, num_heads=2)).eval()
    t5.save_pretrained(os.path.join(root, "text_encoder_3"))

    loaded = load_text_encoders(True, root, torch.device("cpu"), {})
    assert loaded is not None
    l_encs, l_toks = loaded
    assert [type(e).__name__ for e in l_encs] == [
        "CLIPTextModelWithProjection", "CLIPTextModelWithProjection", "T5EncoderModel"]
    assert all(not p.requires_grad for e in l_encs for p in e.parameters())
    prompts = ["the car", "a truck at night"]
    got = text_conditions(True, l_encs, l_toks, prompts, 4, 3, "cpu", torch.float32, None, True)
    want = text_conditions(True, [encs[0], encs[1], t5], [tok, tok, t5tok], prompts, 4, 3, "cpu",
                           torch.float32, None, True)
    assert got[0].shape == (4, 4, 3, 154, 96) and got[1].shape == (4, 4, 3, 40)
    assert torch.allclose(got[0], want[0], atol=1e-6) and torch.allclose(got[1], want[1], atol=1e-6)
    assert torch.equal(got[0][0], got[0][1])             # CFG half: both prompts are ""
    # SD-2.1 layout: CLIPTextModel + one tokenizer
    root21 = os.path.join(root, "sd21")
    tok.save_pretrained(os.path.join(root21, "tokenizer"))
    clip21.save_pretrained(os.path.join(root21, "text_encoder"))
    enc21, tok21 = load_text_encoders(False, root21, torch.device("cpu"), {})
    nested = [[["a", "b", "c"]] * 4]
    ehs, pooled = text_conditions(False, enc21, tok21, nested, 4, 3, "cpu", torch.float32)
    assert ehs.shape == (1, 4, 3, 77, 48) and pooled is None
    assert flatten_prompts(nested)[1] == [1, 4, 3]
    with pytest.raises(RuntimeError, match="no text encoders were loaded"):
        from dwm.models.crossview_temporal_dit import DiTCrossviewTemporalConditionModel
        from dwm.pipelines.ctsd import CrossviewTemporalSD
        CrossviewTemporalSD.get_conditions(
            object.__new__(DiTCrossviewTemporalConditionModel), "pre-encoded", "pre-encoded", {},
            (1, 4, 3, 4, 2, 3), {"pts": torch.zeros(1, 4, 3), "clip_text": ["x"]}, "cpu",
            torch.float32)
