from .dictionnary import Dict
from .logger import set_logging, display_info, get_date, PlotDrawer, export_parameters_to_json
from .parser import parse
from .data_tool import collate_embedding, simple_collate, load_dataset_of_embeddings, Corrupter, Normalizer, get_mask_transformation
from .metering import get_rmse, RankingLoss, CombinedCriterion
from .fused_step import FusedStep
from .inference import ComplementarityScorer
from .embedding_file import write_cemb, read_cemb, convert_json_to_cemb, load_cemb_dataset
