"""Host-side mirror of jellyfysh/mediator for the device path: CudaBatchedMediator."""
