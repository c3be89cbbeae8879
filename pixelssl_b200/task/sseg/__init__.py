from . import model, criterion, func
