from . import static  # noqa: F401
from .static import Calculator, StaticCalculator  # noqa: F401
