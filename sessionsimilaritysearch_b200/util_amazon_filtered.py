"""Drop-in module name for the reference's util_amazon_filtered.py: `from util_amazon_filtered import normalize,
get_item, sequence_to_graph, get_query, session_to_text` (test_amazon_filterd.py:23,30) keeps working when this
package directory is on sys.path.  normalize runs on the GPU (index.normalize); the rest is host featurisation."""
from .index import normalize  # noqa: F401
from .sessions import (get_all_query, get_item, get_item_pos_cnt, get_item_title, get_item_type,  # noqa: F401
                       get_next_query, get_query, get_query_node_tokens, get_session_item_title, sequence_to_graph,
                       session_to_text)
