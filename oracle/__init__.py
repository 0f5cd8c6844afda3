"""CPU oracle for the Gram + attention head: TEST INFRASTRUCTURE, never imported by the product package."""
