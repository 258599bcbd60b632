"""Configuration classes with the reference's attribute surface (nested plain classes, class
attributes) for the robots / controllers / envs / sensors / tasks the two hot paths serve.
Values restate aerial_gym/config/** of the reference (cited per module)."""
import os

PACKAGE_DIRECTORY = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RESOURCES_DIRECTORY = os.path.join(PACKAGE_DIRECTORY, "resources")
