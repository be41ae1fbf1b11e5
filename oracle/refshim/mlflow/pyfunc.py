"""Stand-in for ``mlflow.pyfunc`` base classes (TEST INFRASTRUCTURE)."""


class PythonModel:
    pass


class PythonModelContext:
    artifacts: dict = {}
