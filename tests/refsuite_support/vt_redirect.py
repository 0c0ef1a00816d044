"""pytest plugin: makes the reference's `utils.video_segmenter`, `utils.video_utils` and `utils.budget_planner`
resolve to this repo's modules, the way INTEGRATION.md section 1 prescribes, so that the reference's own test files run
against the drop-in unchanged."""
import sys

import video_transformer_b200.budget_planner as bp
import video_transformer_b200.video_segmenter as vs
import video_transformer_b200.video_utils as vu

sys.modules["utils.video_segmenter"] = vs
sys.modules["utils.video_utils"] = vu
sys.modules["utils.budget_planner"] = bp
