from .. import AbsmaxObserver, AbsmaxObserverLayer  # noqa: F401
