from .basemodel import BaseModel, Linear
from .xdeepfm import xDeepFM
