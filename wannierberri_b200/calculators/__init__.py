from . import static  # noqa: F401
from . import dynamic  # noqa: F401
from .static import Calculator, StaticCalculator  # noqa: F401
from .dynamic import DynamicCalculator  # noqa: F401
